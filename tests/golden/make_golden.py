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

if __name__ == "__main__":
    chi2 = synth.chi2_table()
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
