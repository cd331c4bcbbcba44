import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ov_plane_b200 import api, synth
S = synth.make_scenario("tiny_points")
ctx = api.Context(S.options, device=0, max_state=128, max_meas_rows=4096, debug=True)
names = ["0 smem broadcast (production)", "1 = 0 without panel/pivinv stores", "2 chain only via smem", "3 shuffle pivot + smem updates",
         "4 shuffle pivot chain only", "5 = 3 with double2 reads", "6 shuffle pivot + eager next entry", "7 = 6 software-pipelined", "8 pre-shuffled operands, fused shift (production)"]
for v, nm in enumerate(names):
    out = np.zeros(16)
    ctx._ck(ctx.lib.ovp_debug_potrf_variants(ctx.h, v, 4, out.ctypes.data_as(C.c_void_p)))
    print("variant %-40s cycles per 16 columns (4 reps): %s  -> %.0f / column warm" % (nm, out[:4].astype(int), out[3] / 16))
for mode, nm in enumerate(["sync mid, warp==0 predicates", "sync end, warp==0 predicates", "sync both, warp==0 predicates", "sync mid, simple predicates", "sync end, simple predicates"]):
    for (nt, nch) in ((32, 1), (256, 3)):
        out = np.zeros(16)
        ctx._ck(ctx.lib.ovp_debug_potrf_cond(ctx.h, nt, nch, 110000, 4, out.ctypes.data_as(C.c_void_p), mode))
        print("chain [%s], %3d threads, %d chain warps: cycles per 16 columns %s" % (nm, nt, nch, out[:4].astype(int)))
