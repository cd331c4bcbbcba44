import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ov_plane_b200 import api, synth
chi2 = synth.chi2_table()
def mk(seed):
    S = synth.make_scenario("cfg3_n512_f600_p8", seed=seed)
    c = api.Context(S.options, device=0, max_state=576, max_meas_rows=40000)
    c.set_chi2_table(chi2)
    ch = synth.load_scenario_into(c, S)
    c.snapshot()
    return S, c, synth.feature_batch(S, ch)
for seed in (102, 103, 104, 105, 106):
    S, c, b = mk(seed)
    for graphs in (1,):
        c.lib.ovp_set_use_graphs(c.h, graphs)
        try:
            for it in range(4):
                c.restore()
                o = c.msckf_update(b, 1.0, 1.0)
            print("seed", seed, "graphs", graphs, "ok planes", o["plane_status"].tolist(), "acc", int((o["feat_status"] == 1).sum()))
        except Exception as e:
            print("seed", seed, "graphs", graphs, "FAIL at it", it, e)
    c.close()
for n in (8,):
    cs = [mk(100 + i) for i in range(n)]
    try:
        for _, c, b in cs:
            c.msckf_prepare(b, 1.0, 1.0)
        for it in range(5):
            for _, c, b in cs:
                c.restore()
                c.msckf_launch()
        for _, c, b in cs:
            c.synchronize()
            o = c.msckf_finish()
        print(n, "concurrent same-seed ctxs OK", o["plane_status"].tolist())
    except Exception as e:
        print(n, "concurrent FAIL", e)
    for _, c, b in cs:
        c.close()
