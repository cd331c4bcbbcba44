"""Profiled plane RANSAC + refinement (ovp_plane_fitting / ovp_optimize_plane) for ncu: the 8 planes of the cfg3 workload in one batch each,
warm-up outside the profiler range (use `ncu --profile-from-start off`).  Run on the GPU box through gpurun."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import planefit_cases  # noqa: E402
from ov_plane_b200 import api, synth  # noqa: E402

S = synth.make_scenario("cfg3_n512_f600_p8", seed=0)
ctx = api.Context(S.options, device=0, max_state=576, max_meas_rows=40000)
ctx.set_chi2_table(synth.chi2_table())
ch = synth.load_scenario_into(ctx, S)
fo, pts = planefit_cases.plane_point_sets(S, seed=0)
pr = planefit_cases.refine_problem(S, ch, seed=0, consistent=True, noise=0.006)
fx = np.zeros(len(pr["feat_offset"]) - 1, dtype=np.int32)


def once():
    st = ctx.plane_fitting(fo, pts, 5, 200.0)[0]
    sr = ctx.optimize_plane(pr["feat_offset"], pr["meas_offset"], pr["meas_clone"], pr["uv_norm"], pr["p_FinG"], pr["cp_inG"], fx, 1.0 / 458.0, 0.01)
    return st, sr


for _ in range(2):
    once()
ctx.synchronize()
torch.cuda.profiler.start()
st, sr = once()
ctx.synchronize()
torch.cuda.profiler.stop()
print("planes fitted", int(st.sum()), "of", len(st), "| refinement converged", int((sr[4][:, 0] == 1).sum()), "iterations", sr[4][:, 1].astype(int).tolist())
