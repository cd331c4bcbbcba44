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

if __name__ == "__main__":
    chi2 = synth.chi2_table()
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
