"""Anchored landmark representations through the C ABI (csrc/anchors.cu) against the CPU oracle: get_feature_jacobian_full with a
representation (device bearing rows + chain-rule kernel), perform_anchor_change and change_anchors (UpdaterSLAM.cpp:684-850)."""
import numpy as np
import pytest

from conftest import make_pair
from ov_plane_b200 import api, synth
from test_cpu_anchors import _anchored_landmark, global_to_anchor, to_lambda

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("do_fej", [0, 1])
@pytest.mark.parametrize("name", ["tiny_points", "cfg2_n256_f200"])
def test_feature_jacobian_full_rep_matches_the_oracle(name, do_fej, chi2_table):
    S = synth.make_scenario(name, seed=1)
    S.options = dict(S.options, do_fej=do_fej)
    ctx, orc, chg, cho = make_pair(S, chi2_table)
    assert list(chg) == list(cho)
    worst = 0.0
    for f in (0, 3, 7):
        a, b = S.meas_offset[f], S.meas_offset[f + 1]
        track = [chg[i] for i in S.meas_clone_idx[a:b]]
        outside = [h for h in chg if h not in track]
        calib = orc.var_get(orc.handle_calib())[0]
        for rep in range(6):
            for anchor in ([track[1]] + outside[:1] if rep >= 2 else [-1]):
                pF = global_to_anchor(orc.var_get(anchor)[0], calib, S.p_FinG[f]) if rep >= 2 else S.p_FinG[f]
                pFf = pF + (0.01 if rep < 2 else 0.0)
                g = ctx.feature_jacobian_full_rep(track, S.uv[a:b], rep, anchor, pF, pFf, 1.0)
                o = orc.feature_jacobian_full_rep(track, S.uv[a:b], rep, anchor, pF, pFf, 1.0)
                assert g[3] == o[3], (rep, anchor, g[3], o[3])                     # x_order as handles: identical
                assert g[0].shape == o[0].shape == (2 * len(track), 1 if rep == 5 else 3) and g[1].shape == o[1].shape
                for x, y in zip(g[:3], o[:3]):
                    worst = max(worst, np.abs(x - y).max() / max(1.0, np.abs(y).max()))
    print("%s do_fej %d: 6 representations x anchor inside / outside the track, max rel diff of H_f, H_x, res vs oracle %.2e" % (name, do_fej, worst))
    assert worst < 1e-11
    ctx.close()


@pytest.mark.parametrize("do_fej", [0, 1])
@pytest.mark.parametrize("rep", [2, 3, 4])
def test_anchor_change_matches_the_oracle(rep, do_fej, chi2_table):
    S = synth.make_scenario("small_planes", seed=1)
    S.options = dict(S.options, do_fej=do_fej, max_clone_size=S.cfg["n_clones"] - 1)
    ctx, orc, chg, cho = make_pair(S, chi2_table)
    for be, ch in ((ctx, chg), (orc, cho)):
        fid, hl, _ = _anchored_landmark(be, S, ch, rep, ch[5])
        be.slam_perform_anchor_change(fid, ch[9])
        assert be.slam_get_representation(fid) == (rep, ch[9])
    Pg, Po = ctx.cov(), orc.cov()
    e1 = np.linalg.norm(Pg - Po) / np.linalg.norm(Po)
    vg, vo = ctx.var_get(ctx.slam_handle(fid)), orc.var_get(orc.slam_handle(fid))
    dv = max(np.abs(vg[0][:3] - vo[0][:3]).max(), np.abs(vg[1][:3] - vo[1][:3]).max())
    # change_anchors: the window is over its limit, the landmark is moved into the oldest clone first, then re-anchored at the state time
    for be, ch in ((ctx, chg), (orc, cho)):
        be.slam_perform_anchor_change(fid, ch[0])
        assert be.slam_change_anchors() == 1
        assert be.slam_get_representation(fid) == (rep, ch[-1])
        assert be.slam_change_anchors() == 0
    Pg, Po = ctx.cov(), orc.cov()
    e2 = np.linalg.norm(Pg - Po) / np.linalg.norm(Po)
    vg, vo = ctx.var_get(ctx.slam_handle(fid)), orc.var_get(orc.slam_handle(fid))
    dv = max(dv, np.abs(vg[0][:3] - vo[0][:3]).max(), np.abs(vg[1][:3] - vo[1][:3]).max())
    print("rep %d do_fej %d: cov rel err after perform_anchor_change %.2e, after change_anchors %.2e, landmark value / fej max diff %.2e" % (
        rep, do_fej, e1, e2, dv))
    assert e1 < 1e-9 and e2 < 1e-9 and dv < 1e-10
    # a clone that still anchors a landmark cannot be marginalised (the reference asserts change_anchors ran first, UpdaterSLAM.cpp:699)
    with pytest.raises(api.OvpError):
        ctx.marginalize(chg[-1])
    # the fused GLOBAL_3D update path refuses the anchored landmark instead of misreading its value
    with pytest.raises(api.OvpError):
        ctx.slam_update(dict(F=1, meas_offset=np.array([0, 2], dtype=np.int32), meas_clone=np.array(chg[:2], dtype=np.int32),
                             uv=np.zeros((2, 2), dtype=np.float32), featid=np.array([fid], dtype=np.int64), planeid=np.zeros(1, dtype=np.int64)))
    ctx.close()
