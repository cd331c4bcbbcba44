"""GPU parity of the SLAM-landmark half of the path: UpdaterSLAM::delayed_init and UpdaterSLAM::update
(update/UpdaterSLAM.cpp:225-372, :389-735) against the CPU restatement, through the C ABI."""
import numpy as np
import pytest

from conftest import make_pair
from ov_plane_b200 import synth
from test_gpu_parity import relerr

pytestmark = pytest.mark.gpu


def _values_close(ctx, orc, hg, ho):
    vg, fg = ctx.var_get(int(hg))
    vo, fo = orc.var_get(int(ho))
    assert np.allclose(vg, vo, rtol=1e-6, atol=1e-7) and np.allclose(fg, fo, rtol=1e-9, atol=1e-12)


def _slam_case(name, seed, chi2_table, nslam, perturb=0.0):
    S = synth.make_scenario(name, seed=seed)
    ctx, orc, chg, cho = make_pair(S, chi2_table, max_state=S.N + 3 * nslam + 64)
    sel = np.arange(min(nslam, S.F))
    if perturb:  # attach some features to the WRONG in-state plane so that the plane -> no-plane fallback is exercised
        rng = np.random.default_rng(seed + 99)
        ids = [p[0] for p in S.planes]
        for f in rng.choice(len(sel), size=max(1, len(sel) // 4), replace=False):
            others = [i for i in ids if i != S.planeid[f]]
            if others:
                S.planeid[f] = others[int(rng.integers(len(others)))]
    bg, bo = synth.feature_batch(S, chg, sel), synth.feature_batch(S, cho, sel)
    return S, ctx, orc, bg, bo, chg, cho


@pytest.mark.parametrize("name,seed,nslam,perturb", [("tiny_planes", 0, 6, 0.0), ("small_planes", 0, 40, 0.0), ("small_planes", 1, 40, 1.0), ("small_planes", 3, 40, 1.0),
                                                     ("tiny_points", 0, 8, 0.0), ("cfg3_n512_f600_p8", 0, 20, 1.0)])
def test_slam_delayed_init_then_update(name, seed, nslam, perturb, chi2_table):
    S, ctx, orc, bg, bo, chg, cho = _slam_case(name, seed, chi2_table, nslam, perturb)
    n0 = ctx.cov_rows()
    g = ctx.slam_delayed_init(bg, 1.0, 1.0)
    o = orc.slam_delayed_init(bo, 1.0, 1.0)
    print(name, seed, "delayed_init status gpu", np.bincount(g["feat_status"], minlength=4), "ref", np.bincount(o["feat_status"], minlength=4),
          "N", n0, "->", ctx.cov_rows())
    assert np.array_equal(g["feat_status"], o["feat_status"])
    assert ctx.cov_rows() == orc.cov_rows()
    assert (g["feat_status"] > 0).any()
    for hg, ho, fid in zip(g["new_handles"], o["new_handles"], bg["featid"]):
        assert (hg >= 0) == (ho >= 0)
        if hg >= 0:
            assert ctx.slam_handle(fid) == hg and orc.slam_handle(fid) == ho
            assert ctx.var_id(int(hg)) == orc.var_id(int(ho))
            _values_close(ctx, orc, hg, ho)
    e = relerr(ctx.cov(), orc.cov())
    print("cov rel err after delayed_init %.3e" % e)
    assert e < 1e-6

    # ---- UpdaterSLAM::update on the landmarks that made it into the state (same tracks: a second look at the same pixels) ----
    keep = np.nonzero(g["feat_status"] > 0)[0]
    ug, uo = synth.feature_batch(S, chg, keep), synth.feature_batch(S, cho, keep)
    rng = np.random.default_rng(seed + 7)
    noise = rng.normal(0.0, 0.5, ug["uv"].shape).astype(np.float32)
    ug["uv"] = ug["uv"] + noise
    uo["uv"] = uo["uv"] + noise
    big = rng.choice(len(keep), size=max(1, len(keep) // 8), replace=False)  # a few gross outliers -> rejected / should_marg
    for b in (ug, uo):
        for f in big:
            b["uv"][b["meas_offset"][f]:b["meas_offset"][f + 1]] += 25.0
    g2 = ctx.slam_update(ug, 1.0, 1.0)
    o2 = orc.slam_update(uo, 1.0, 1.0)
    print("slam_update status gpu", np.bincount(g2["feat_status"], minlength=4), "ref", np.bincount(o2["feat_status"], minlength=4))
    assert np.allclose(g2["feat_chi2"], o2["feat_chi2"], rtol=1e-6, atol=1e-8)
    assert np.array_equal(g2["feat_status"], o2["feat_status"])
    for fid in ug["featid"]:
        assert ctx.slam_should_marg(fid) == orc.slam_should_marg(fid)
    e = relerr(ctx.cov(), orc.cov())
    print("cov rel err after slam_update %.3e" % e)
    assert e < 1e-6
    for hg, ho in zip(g["new_handles"][keep], o["new_handles"][keep]):
        _values_close(ctx, orc, hg, ho)
    vg, vo = ctx.var_get(ctx.handle_imu())[0], orc.var_get(orc.handle_imu())[0]
    assert np.allclose(vg, vo, rtol=1e-7, atol=1e-9)
    # marginalize_slam removes exactly the flagged landmarks on both sides (StateHelper.cpp:638-652)
    ctx.marginalize_slam()
    orc.marginalize_slam()
    assert ctx.cov_rows() == orc.cov_rows()
    assert relerr(ctx.cov(), orc.cov()) < 1e-6


def test_slam_update_without_plane_constraint(chi2_table):
    S, ctx, orc, bg, bo, chg, cho = _slam_case("small_planes", 2, chi2_table, 30)
    g = ctx.slam_delayed_init(bg, 1.0, 1.0, use_plane_constraint=False)
    assert (g["feat_status"] != 3).all()
    keep = np.nonzero(g["feat_status"] > 0)[0]
    ug = synth.feature_batch(S, chg, keep)
    g2 = ctx.slam_update(ug, 1.0, 1.0, use_plane_constraint=False)
    assert (g2["feat_status"] != 3).all() and (g2["feat_status"] == 1).any()
    P = ctx.cov()
    assert np.allclose(P, P.T, rtol=0, atol=1e-12 * np.abs(P).max()) and np.linalg.eigvalsh(P).min() > -1e-12


@pytest.mark.parametrize("name,seed,nslam", [("tiny_planes", 0, 6), ("tiny_points", 0, 8)])
def test_slam_against_committed_golden_vectors(name, seed, nslam, chi2_table):
    """the CUDA path against the committed fixtures (tests/golden/slam_*.npz, written by the oracle): no oracle at run time"""
    import os
    import sys
    from ov_plane_b200 import api
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gold)
    import make_golden
    g = np.load(os.path.join(gold, "slam_%s_s%d.npz" % (name, seed)))
    r = make_golden.slam_case(name, seed, nslam, lambda S: api.Context(S.options, device=0, max_state=S.N + 3 * nslam + 64, max_meas_rows=8192),
                              chi2_table)
    assert np.array_equal(r["init_status"], g["init_status"]) and np.array_equal(r["upd_status"], g["upd_status"])
    assert relerr(r["P_init"], g["P_init"]) < 1e-6 and relerr(r["P_upd"], g["P_upd"]) < 1e-6
    assert np.allclose(r["upd_chi2"], g["upd_chi2"], rtol=1e-6, atol=1e-8)
    assert np.allclose(r["imu"], g["imu"], rtol=1e-7, atol=1e-9)
