"""Generates the committed golden vectors from the CPU oracle (the reference itself cannot be built or imported here:
no Eigen / Boost / ov_core).  They pin the oracle against regressions and give the `-m gpu` tests a fixed
target that does not depend on re-running the oracle.  Usage: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_backend as ob  # noqa: E402
from ov_plane_b200 import synth  # noqa: E402

CASES = [("tiny_points", 0), ("tiny_planes", 0), ("small_planes", 1), ("cfg1_euroc_n96", 0), ("cfg2_n256_f200", 0)]

def slam_case(name, seed, nslam, backend_factory, chi2):
    """delayed_init of the first nslam features, then a SLAM update with perturbed pixels (shared by the golden writer and the tests)"""
    S = synth.make_scenario(name, seed=seed)
    be = backend_factory(S)
    be.set_chi2_table(chi2)
    ch = synth.load_scenario_into(be, S)
    sel = np.arange(min(nslam, S.F))
    d = be.slam_delayed_init(synth.feature_batch(S, ch, sel), 1.0, 1.0)
    P_init = be.cov().copy()
    keep = np.nonzero(d["feat_status"] > 0)[0]
    u = synth.feature_batch(S, ch, keep)
    rng = np.random.default_rng(seed + 7)
    u["uv"] = u["uv"] + rng.normal(0.0, 0.5, u["uv"].shape).astype(np.float32)
    u["uv"][u["meas_offset"][1]:u["meas_offset"][2]] += 25.0  # one gross outlier
    r = be.slam_update(u, 1.0, 1.0)
    return dict(init_status=d["feat_status"], P_init=P_init, upd_status=r["feat_status"], upd_chi2=r["feat_chi2"], P_upd=be.cov(),
                imu=be.var_get(be.handle_imu())[0])


SLAM_CASES = [("tiny_planes", 0, 6), ("tiny_points", 0, 8)]


def planefit_case(name, seed, backend_factory, chi2):
    """PlaneFitting on the planes of a scenario (RANSAC batch with extra degenerate candidates, refinement batch with a mixed fix_plane vector)
    and one anchor change - shared by the golden writer, the CPU regression test and the GPU test."""
    import planefit_cases
    S = synth.make_scenario(name, seed=seed)
    S.options = dict(S.options, max_clone_size=S.cfg["n_clones"] - 1)
    be = backend_factory(S)
    be.set_chi2_table(chi2)
    ch = synth.load_scenario_into(be, S)
    fo, pts = planefit_cases.plane_point_sets(S, seed=seed)
    rng = np.random.RandomState(5)
    for e in (rng.randn(4, 3), np.array([1.0, 2.0, 3.0]) + 0.004 * rng.randn(30, 3)):
        fo = np.append(fo, fo[-1] + len(e)).astype(np.int32)
        pts = np.vstack([pts, e])
    st, ab, inl = be.plane_fitting(fo, pts, 5, 200.0)
    pr = planefit_cases.refine_problem(S, ch, seed=seed, consistent=True, noise=0.006)
    fx = np.zeros(len(pr["feat_offset"]) - 1, dtype=np.int32)
    fx[1::2] = 1
    rs, rp, rc, ri, rinfo = be.optimize_plane(pr["feat_offset"], pr["meas_offset"], pr["meas_clone"], pr["uv_norm"], pr["p_FinG"], pr["cp_inG"], fx,
                                              1.0 / 458.0, 0.01)
    # anchor change of the scenario's landmark: MSCKF inverse depth anchored in clone 5, moved to clone 9, then through change_anchors
    from test_cpu_anchors import _anchored_landmark
    fid, hl, _ = _anchored_landmark(be, S, ch, 4, ch[5])
    be.slam_perform_anchor_change(fid, ch[9])
    be.slam_perform_anchor_change(fid, ch[0])
    n_changed = be.slam_change_anchors()
    v, f = be.var_get(hl)
    return dict(fit_status=st, fit_abcd=ab, fit_inlier=inl, ref_status=rs, ref_p=rp, ref_cp=rc, ref_inlier=ri, ref_info=rinfo,
                anchor_value=np.asarray(v[:3]), anchor_fej=np.asarray(f[:3]), anchor_changed=np.array([n_changed]), P_anchor=be.cov())


PLANEFIT_CASES = [("small_planes", 1), ("cfg3_n512_f600_p8", 0)]

if __name__ == "__main__":
    chi2 = synth.chi2_table()
    for name, seed in PLANEFIT_CASES:
        g = planefit_case(name, seed, lambda S: ob.OracleContext(S.options), chi2)
        np.savez_compressed(os.path.join(HERE, "planefit_%s_s%d.npz" % (name, seed)), **g)
        print("planefit", name, seed, "fitted", g["fit_status"], "refined", g["ref_status"], "iterations", g["ref_info"][:, 1].astype(int), "written")
    if len(sys.argv) > 1 and sys.argv[1] == "planefit":
        sys.exit(0)
    for name, seed, nslam in SLAM_CASES:
        g = slam_case(name, seed, nslam, lambda S: ob.OracleContext(S.options), chi2)
        np.savez_compressed(os.path.join(HERE, "slam_%s_s%d.npz" % (name, seed)), **g)
        print("slam", name, seed, "init", g["init_status"], "update", g["upd_status"], "written")
    for name, seed in CASES:
        S = synth.make_scenario(name, seed=seed)
        o = ob.OracleContext(S.options)
        o.set_chi2_table(chi2)
        ch = synth.load_scenario_into(o, S)
        r = o.msckf_update(synth.feature_batch(S, ch), 1.0, 1.0)
        vals = np.stack([o.var_get(h)[0] for h in ch])
        np.savez_compressed(os.path.join(HERE, "%s_s%d.npz" % (name, seed)), P0=S.P0.astype(np.float64), uv=S.uv,
                            P1=o.cov(), imu1=o.var_get(o.handle_imu())[0], clones1=vals, feat_status=r["feat_status"],
                            feat_chi2=r["feat_chi2"], plane_status=r["plane_status"], plane_chi2=r["plane_chi2"],
                            hx_order_clone_idx=np.array([ch.index(h) if h in ch else -1 - h for h in r["hx_order"]]))
        print(name, seed, "N", S.N, "written")
