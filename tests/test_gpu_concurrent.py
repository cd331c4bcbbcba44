"""Independent filters on one GPU (one context / stream / CUDA graph each) must reproduce their solo results bit for bit:
the fused Cholesky synchronises its CTAs through flags in global memory, and a timing-dependent hazard between them only
shows when several cooperative launches share the GPU (tools/diag_concurrent3.py is the longer stress version)."""
import numpy as np
import pytest

from ov_plane_b200 import api, synth

pytestmark = pytest.mark.gpu


def test_concurrent_filters_reproduce_solo_results(chi2_table):
    C_f, rounds = 6, 12
    ctxs, refs = [], []
    for i in range(C_f):
        S = synth.make_scenario("cfg3_n512_f600_p8", seed=100 + i)
        c = api.Context(S.options, device=0, max_state=576, max_meas_rows=40000)
        c.set_chi2_table(chi2_table)
        ch = synth.load_scenario_into(c, S)
        b = synth.feature_batch(S, ch)
        c.snapshot()
        c.msckf_update(b, 1.0, 1.0)
        refs.append(c.cov().copy())
        c.restore()
        c.msckf_prepare(b, 1.0, 1.0)
        ctxs.append(c)
    bad = []
    for r in range(rounds):
        for c in ctxs:
            c.restore()
            c.msckf_launch()
        for i, c in enumerate(ctxs):
            c.msckf_finish()
            if not np.array_equal(c.cov(), refs[i]):
                bad.append((r, i, float(np.linalg.norm(c.cov() - refs[i]) / np.linalg.norm(refs[i]))))
    for c in ctxs:
        c.close()
    assert not bad, bad[:5]
