import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ov_plane_b200 import api, synth
chi2 = synth.chi2_table()
def mk(seed):
    S = synth.make_scenario("cfg3_n512_f600_p8", seed=seed)
    c = api.Context(S.options, device=0, max_state=576, max_meas_rows=40000)
    c.set_chi2_table(chi2)
    ch = synth.load_scenario_into(c, S)
    c.snapshot()
    c.lib.ovp_set_use_graphs(c.h, GRAPHS)
    return S, c, synth.feature_batch(S, ch)
variant = sys.argv[1] if len(sys.argv) > 1 else "all"
GRAPHS = int(sys.argv[2]) if len(sys.argv) > 2 else 1
S, ctx, batch = mk(0)
for _ in range(3):
    ctx.restore(); ctx.msckf_update(batch, 1.0, 1.0)
if variant in ("all", "prof"):
    ctx.msckf_prepare(batch, 1.0, 1.0)
    ctx.set_profiling(1)
    for _ in range(3):
        ctx.restore(); ctx.msckf_launch()
    ctx.profile_report(); ctx.set_profiling(0); ctx.msckf_finish()
if variant == "eager3":
    ctx.msckf_prepare(batch, 1.0, 1.0)
    ctx.lib.ovp_set_use_graphs(ctx.h, 0)
    for _ in range(3):
        ctx.restore(); ctx.msckf_launch()
    ctx.msckf_finish(); ctx.lib.ovp_set_use_graphs(ctx.h, GRAPHS)
if variant in ("all", "torch"):
    a = torch.randn(4096, 4096, dtype=torch.float64, device="cuda"); b2 = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
    for _ in range(3): torch.matmul(a, b2)
    torch.cuda.synchronize()
if variant in ("all", "self"):
    print("own dgemm", ctx.selftest_dgemm_tflops(2048, 10))
others = [mk(100 + i) for i in range(7)]
for _, c, b in others:
    c.msckf_prepare(b, 1.0, 1.0)
ctx.msckf_prepare(batch, 1.0, 1.0)
allc = [(S, ctx, batch)] + others
for it in range(6):
    for _, c, b in allc:
        c.restore(); c.msckf_launch()
for i, (_, c, b) in enumerate(allc):
    try:
        c.synchronize(); o = c.msckf_finish()
        print(variant, GRAPHS, "ctx", i, "ok", o["plane_status"].tolist())
    except Exception as e:
        print(variant, GRAPHS, "ctx", i, "FAIL", e)
