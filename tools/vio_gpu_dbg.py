"""Debug aid: the VIO loop on the GPU and on the oracle side by side, per-frame differences."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from ov_plane_b200 import api, synth, vio_sim
import oracle_backend as ob
planes = int(sys.argv[1]) if len(sys.argv) > 1 else 1
opts = vio_sim.state_options(max_clones=11)
chi2 = synth.chi2_table()
g = api.Context(opts, device=0, max_state=384, max_meas_rows=20000); g.set_chi2_table(chi2)
o = ob.OracleContext(opts); o.set_chi2_table(chi2)
lg, cg = vio_sim.run(g, n_frames=120, seed=3, max_clones=11, n_feats=60, use_planes=bool(planes), keep_cov_every=1)
lo, co = vio_sim.run(o, n_frames=120, seed=3, max_clones=11, n_feats=60, use_planes=bool(planes), keep_cov_every=1, gate_ctx=lambda: ob.GaugeProbe(gate_without=True))
for k, (rg, ro) in enumerate(zip(lg.frames, lo.frames)):
    dP = np.linalg.norm(cg[k][1] - co[k][1]) / np.linalg.norm(co[k][1])
    same = ("feat_status" not in ro) or (np.array_equal(rg["feat_status"], ro["feat_status"]) and np.array_equal(rg["plane_status"], ro["plane_status"]))
    if k % 4 == 0 or not same:
        print("t %.1f N %d used %s/%s planes %s init %s same_gates %s | imu diff %.2e cov diff %.2e" % (rg["t"], rg["N"], rg.get("n_used"), ro.get("n_used"), rg.get("n_planes"), rg.get("planes_initialised"), same, np.abs(rg["imu"] - ro["imu"]).max(), dP))
