"""Debug aid: per-frame parity of the VIO loop (GPU re-seeded from the oracle every frame); prints the frames with the largest single-frame difference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from ov_plane_b200 import synth
import test_gpu_vio as T
chi2 = synth.chi2_table()
for clones, frames in ((11, 130), (30, 120)):
    g, o, lg, lo = T._pair(clones, chi2, seed=3)
    rows = []
    for k in range(1, frames + 1):
        t = 0.5 + k * lg.sim.cam_dt
        recs = []
        for lp in (lg, lo):
            lp._feed_imu(t)
            recs.append(lp.step(t, lp.sim.camera_frame(t)))
        rg, ro = recs
        d = np.abs(rg["imu"] - ro["imu"])
        dP = np.linalg.norm(g.cov() - o.cov()) / np.linalg.norm(o.cov())
        rows.append((d.max(), t, d, dP, rg.get("n_msckf"), rg.get("n_used"), rg.get("n_planes"), rg.get("planes_initialised"), rg.get("plane_status")))
        T._copy_state(lo, lg)
    print("window", clones)
    for r in sorted(rows, key=lambda x: -max(x[0], x[3]))[:6]:
        print("t %.1f max diff %.2e cov %.2e | q %.1e p %.1e v %.1e bg %.1e ba %.1e | msckf %s used %s planes %s init %s pstatus %s" % (
            r[1], r[0], r[3], r[2][:4].max(), r[2][4:7].max(), r[2][7:10].max(), r[2][10:13].max(), r[2][13:16].max(), r[4], r[5], r[6], r[7], r[8]))
