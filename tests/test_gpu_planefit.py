"""PlaneFitting on the device (ovp_plane_fitting, ovp_optimize_plane: csrc/planefit.cu) against the CPU oracle (oracle/oracle_planefit.hpp).
The RANSAC decisions (status, inlier sets) are index-exact; plane parameters and refined positions agree to 1e-9 / 1e-7 relative.  The oracle
solves the dogleg iteration with dense normal equations on Jacobian rows, the device through per-feature blocks and a Schur complement - two
formulations of the same restated Ceres algorithm (parity against Ceres itself is unpinned: it is not available here)."""
import numpy as np
import pytest

import planefit_cases
from conftest import make_pair
from ov_plane_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("name,seed", [("tiny_planes", 0), ("small_planes", 0), ("small_planes", 1), ("cfg3_n512_f600_p8", 0)])
def test_plane_fitting_matches_the_oracle(name, seed, kind, chi2_table):
    S = synth.make_scenario(name, seed=seed)
    ctx, orc, chg, cho = make_pair(S, chi2_table)
    fo, pts = planefit_cases.plane_point_sets(S, seed=seed)
    rng = np.random.RandomState(5)
    # extra candidate planes in the same batch: too few points, a cluster without five separated points, a cloud that is no plane
    extra = [rng.randn(4, 3), np.array([1.0, 2.0, 3.0]) + 0.004 * rng.randn(30, 3), rng.uniform(-2, 2, size=(40, 3))]
    for e in extra:
        fo = np.append(fo, fo[-1] + len(e)).astype(np.int32)
        pts = np.vstack([pts, e])
    sg, ag, ig = ctx.plane_fitting(fo, pts, 5, 200.0, shuffle_kind=kind)
    so, ao, io = orc.plane_fitting(fo, pts, 5, 200.0, shuffle_kind=kind)
    assert np.array_equal(sg, so), (sg, so)
    assert np.array_equal(ig, io), np.nonzero(ig != io)
    assert (sg[-3:] == 0).all() and sg[:-3].sum() >= len(sg) - 4
    d = np.abs(ag - ao).max()
    print("%s seed %d shuffle %d: %d planes fitted of %d candidates, inliers %d of %d points, max |abcd gpu - oracle| %.2e" % (
        name, seed, kind, sg.sum(), len(sg), ig.sum(), len(ig), d))
    assert d < 1e-9
    ctx.close()


@pytest.mark.parametrize("name,seed,consistent,fix", [("small_planes", 2, True, 0), ("small_planes", 2, True, 1), ("small_planes", 1, False, 0),
                                                      ("tiny_planes", 0, True, 0), ("cfg3_n512_f600_p8", 0, True, 0), ("cfg3_n512_f600_p8", 0, True, 1),
                                                      ("cfg3_n512_f600_p8", 1, False, 1)])
def test_optimize_plane_matches_the_oracle(name, seed, consistent, fix, chi2_table):
    S = synth.make_scenario(name, seed=seed)
    ctx, orc, chg, cho = make_pair(S, chi2_table)
    pr = planefit_cases.refine_problem(S, chg, seed=seed, consistent=consistent, noise=0.006)
    assert np.array_equal(np.asarray(chg), np.asarray(cho))
    fx = np.full(len(pr["feat_offset"]) - 1, fix, dtype=np.int32)
    if len(fx) > 2:
        fx[1] = 1 - fix  # mixed batch
    args = (pr["feat_offset"], pr["meas_offset"], pr["meas_clone"], pr["uv_norm"], pr["p_FinG"], pr["cp_inG"], fx, 1.0 / 458.0, 0.01)
    sg, pg, cg, ig, ng = ctx.optimize_plane(*args)
    so, po, co, io, no = orc.optimize_plane(*args)
    print("%s seed %d consistent %d: status gpu %s oracle %s | iterations gpu %s oracle %s | reason %s | cost %s -> gpu %s oracle %s" % (
        name, seed, consistent, sg, so, ng[:, 1].astype(int), no[:, 1].astype(int), no[:, 4].astype(int), np.round(no[:, 2], 3), np.round(ng[:, 3], 6),
        np.round(no[:, 3], 6)))
    assert np.array_equal(ng[:, 0], no[:, 0]) and np.array_equal(ng[:, 1], no[:, 1]) and np.array_equal(ng[:, 4], no[:, 4])
    assert np.allclose(ng[:, 2], no[:, 2], rtol=1e-10, atol=0) and np.allclose(ng[:, 3], no[:, 3], rtol=1e-8, atol=1e-12)
    assert np.array_equal(sg, so) and np.array_equal(ig, io)
    dp, dc = np.abs(pg - po).max(), np.abs(cg - co).max()
    print("   max |p gpu - oracle| %.2e m, max |cp gpu - oracle| %.2e m, inliers %d of %d" % (dp, dc, ig.sum(), len(ig)))
    assert dp < 1e-7 and dc < 1e-7
    if consistent:
        assert (no[:, 0] == 1).sum() >= 1 or name == "tiny_planes"  # (few short tracks: neither side reaches CONVERGENCE in 12 iterations)
    else:
        assert (no[:, 0] == 0).all() and np.array_equal(pg, pr["p_FinG"]) and np.array_equal(cg, pr["cp_inG"])  # no CONVERGENCE: untouched
    ctx.close()


def test_optimize_plane_edge_cases(chi2_table):
    S = synth.make_scenario("small_planes", seed=2)
    ctx, orc, chg, cho = make_pair(S, chi2_table)
    pr = planefit_cases.refine_problem(S, chg, seed=2, consistent=True, noise=0.0, px_noise=0.0, slam_share=0.0)
    fx = np.zeros(len(pr["feat_offset"]) - 1, dtype=np.int32)
    args = (pr["feat_offset"], pr["meas_offset"], pr["meas_clone"], pr["uv_norm"], pr["p_FinG"], pr["cp_inG"], fx, 1.0 / 458.0, 0.01)
    sg, pg, cg, ig, ng = ctx.optimize_plane(*args)
    so, po, co, io, no = orc.optimize_plane(*args)
    assert (sg == 1).all() and np.array_equal(sg, so) and np.array_equal(ng[:, 1], no[:, 1]) and ig.all()
    assert np.abs(cg - co).max() < 1e-8 and (ng[:, 3] < 1e-6).all()
    # tracks longer than a warp (every measurement listed three times: up to 36 per feature): the residual pass strides the lanes over the measurements
    mo2 = (3 * pr["meas_offset"]).astype(np.int32)
    mc2 = np.concatenate([np.tile(pr["meas_clone"][a:b], 3) for a, b in zip(pr["meas_offset"][:-1], pr["meas_offset"][1:])]).astype(np.int32)
    uv2 = np.concatenate([np.tile(pr["uv_norm"][a:b], (3, 1)) for a, b in zip(pr["meas_offset"][:-1], pr["meas_offset"][1:])]).astype(np.float32)
    noisy = pr["p_FinG"] + 0.004 * np.random.RandomState(3).randn(*pr["p_FinG"].shape)
    a3 = (pr["feat_offset"], mo2, mc2, uv2, noisy, pr["cp_inG"], fx, 1.0 / 458.0, 0.01)
    sg, pg, cg, ig, ng = ctx.optimize_plane(*a3)
    so, po, co, io, no = orc.optimize_plane(*a3)
    assert int(np.diff(mo2).max()) > 32
    assert np.array_equal(sg, so) and np.array_equal(ig, io) and np.array_equal(ng[:, 1], no[:, 1]) and np.array_equal(ng[:, 4], no[:, 4])
    assert np.abs(pg - po).max() < 1e-7 and np.abs(cg - co).max() < 1e-7
    # one SLAM-only feature against a fixed plane (UpdaterSLAM.cpp:171 style), three features with a free plane (refused), an empty candidate
    fo = np.array([0, 1, 4, 4], dtype=np.int32)
    mo = np.concatenate([[0, 0], pr["meas_offset"][1:4]]).astype(np.int32)
    p0 = pr["p_FinG"][:4]
    cp = pr["cp_inG"][[0, 0, 0]]
    a2 = (fo, mo, pr["meas_clone"], pr["uv_norm"], p0, cp, np.array([1, 0, 0], dtype=np.int32), 1.0 / 458.0, 0.01)
    sg, pg, cg, ig, ng = ctx.optimize_plane(*a2)
    so, po, co, io, no = orc.optimize_plane(*a2)
    assert np.array_equal(sg, so) and np.array_equal(ig, io) and np.array_equal(ng[:, 4], no[:, 4]), (sg, so, ng, no)
    assert ng[0, 4] == 4 and sg[1] == 0 and sg[2] == 0
    ctx.close()


@pytest.mark.parametrize("name,seed", [("small_planes", 1), ("cfg3_n512_f600_p8", 0)])
def test_planefit_and_anchor_change_against_committed_golden_vectors(name, seed, chi2_table):
    """the CUDA path against the committed fixtures (tests/golden/planefit_*.npz, written by the oracle): no oracle at run time"""
    import os
    import sys
    from ov_plane_b200 import api
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gold)
    import make_golden
    g = np.load(os.path.join(gold, "planefit_%s_s%d.npz" % (name, seed)))
    r = make_golden.planefit_case(name, seed, lambda S: api.Context(S.options, device=0, max_state=S.N + 64, max_meas_rows=60000), chi2_table)
    for k in ("fit_status", "fit_inlier", "ref_status", "ref_inlier", "anchor_changed"):
        assert np.array_equal(r[k], g[k]), k
    assert np.array_equal(r["ref_info"][:, [0, 1, 4]], g["ref_info"][:, [0, 1, 4]])   # converged, iterations, termination reason
    assert np.abs(r["fit_abcd"] - g["fit_abcd"]).max() < 1e-9
    assert np.abs(r["ref_p"] - g["ref_p"]).max() < 1e-7 and np.abs(r["ref_cp"] - g["ref_cp"]).max() < 1e-7
    assert np.abs(r["anchor_value"] - g["anchor_value"]).max() < 1e-10 and np.abs(r["anchor_fej"] - g["anchor_fej"]).max() < 1e-10
    assert np.linalg.norm(r["P_anchor"] - g["P_anchor"]) / np.linalg.norm(g["P_anchor"]) < 1e-9
