"""Stress: C independent filters on one GPU (separate streams / contexts), R rounds; every filter must reproduce its solo result."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ov_plane_b200 import api, synth
C_f, R = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 30
chi2 = synth.chi2_table()
ctxs, batches, refs = [], [], []
for i in range(C_f):
    S = synth.make_scenario("cfg3_n512_f600_p8", seed=100 + i)
    c = api.Context(S.options, device=0, max_state=576, max_meas_rows=40000)
    c.set_chi2_table(chi2)
    ch = synth.load_scenario_into(c, S)
    b = synth.feature_batch(S, ch)
    c.snapshot()
    c.msckf_update(b, 1.0, 1.0)
    refs.append(c.cov().copy())
    c.restore()
    c.msckf_prepare(b, 1.0, 1.0)
    ctxs.append(c); batches.append(b)
bad = 0
for r in range(R):
    for c in ctxs:
        c.restore()
        c.msckf_launch()
    for i, c in enumerate(ctxs):
        try:
            c.msckf_finish()
            e = np.linalg.norm(c.cov() - refs[i]) / np.linalg.norm(refs[i])
            if e > 1e-12:
                bad += 1
                print("round", r, "filter", i, "differs from its solo result: rel", e)
        except Exception as ex:
            bad += 1
            print("round", r, "filter", i, "error", ex)
            c.restore()
            c.msckf_prepare(batches[i], 1.0, 1.0)
print("concurrent stress: %d filters x %d rounds, %d bad" % (C_f, R, bad))
