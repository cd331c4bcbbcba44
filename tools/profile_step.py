"""One profiled MSCKF+plane update step for ncu: warm-up outside the profiler range, then `--steps` steps inside
cudaProfilerStart/Stop (use `ncu --profile-from-start off`).  Run on the GPU box through gpurun."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from ov_plane_b200 import api, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg3_n512_f600_p8")
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
S = synth.make_scenario(a.workload, seed=0)
ctx = api.Context(S.options, device=0, max_state=576, max_meas_rows=40000)
ctx.set_chi2_table(synth.chi2_table())
ch = synth.load_scenario_into(ctx, S)
batch = synth.feature_batch(S, ch)
ctx.snapshot()
for _ in range(2):
    ctx.restore()
    ctx.msckf_update(batch, 1.0, 1.0)
ctx.msckf_prepare(batch, 1.0, 1.0)
ctx.synchronize()
torch.cuda.profiler.start()
for _ in range(a.steps):
    ctx.restore()
    ctx.msckf_launch()
ctx.synchronize()
torch.cuda.profiler.stop()
out = ctx.msckf_finish()
print("accepted", int((out["feat_status"] == 1).sum()), "planes", out["plane_status"].tolist())
