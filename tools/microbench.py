import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ov_plane_b200 import api, synth
S = synth.make_scenario("tiny_points")
ctx = api.Context(S.options, device=0, max_state=576, max_meas_rows=4096, debug=True)
for n in (512, 1024, 2048, 4096):
    print("own DMMA gemm n=%d: %.2f TFLOP/s" % (n, ctx.selftest_dgemm_tflops(n, 5)))
lat = np.zeros(16)
ctx._ck(ctx.lib.ovp_debug_fp64_latency(ctx.h, lat.ctypes.data_as(C.c_void_p)))
print("dependent-chain cycles/op: dfma %.1f  rsqrt(double)+add %.1f  1/x+add %.1f  sqrt+add %.1f  shfl(double) %.1f  lds chain %.1f  rsqrtf+2 Newton %.1f" % tuple(lat[:7]))
print("DMMA m8n8k4 one warp: dependent chain %.1f cycles/op, 8 independent accumulators %.1f cycles/op" % (lat[8], lat[9]))
