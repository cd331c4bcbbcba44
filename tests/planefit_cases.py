"""TEST INFRASTRUCTURE: plane-fitting / plane-refinement problems cut out of the synthetic scenarios (shared by the CPU and GPU tests)."""
import numpy as np

from ov_plane_b200 import jpl, synth, vio_sim


def plane_point_sets(S, seed=0, outlier_frac=0.15, noise=0.004):
    """Per in-scenario plane: the (noisy) positions of its features plus a share of off-plane outliers.  Returns feat_offset, points."""
    rng = np.random.RandomState(seed)
    offs, pts = [0], []
    for pid in S.plane_ids:
        idx = np.nonzero(S.planeid == pid)[0]
        p = S.pf_true[idx] + noise * rng.randn(len(idx), 3)
        n_out = int(outlier_frac * len(idx))
        if n_out:
            sel = rng.choice(len(idx), n_out, replace=False)
            p[sel] += rng.uniform(0.08, 0.4, size=(n_out, 1)) * rng.choice([-1.0, 1.0], size=(n_out, 1)) * _normal_of(S, pid)
        pts.append(p)
        offs.append(offs[-1] + len(idx))
    return np.array(offs, dtype=np.int32), np.ascontiguousarray(np.vstack(pts))


def _normal_of(S, pid):
    cp = np.asarray(S.plane_cp)[list(S.plane_ids).index(pid)]
    return cp / np.linalg.norm(cp)


def refine_problem(S, clone_handles, seed=0, slam_share=0.1, drop_planes=(), noise=0.012, consistent=False, px_noise=1.0):
    """The optimize_plane inputs for every plane of a scenario: features grouped by plane, their measurements as undistorted normalised
    coordinates against clone handles, the (perturbed) triangulated positions and plane estimates; a share of the features carries no
    measurements (SLAM features, PlaneFitting.cpp:274-277).  consistent = False: the scenario's own pixels (taken through the TRUE poses while the
    state holds perturbed estimates: residuals of several pixels, a hard robust problem); True: pixels re-projected through the poses the state
    holds plus px_noise pixels (a filter whose poses are good: the regime in which the reference's 12 iterations suffice)."""
    rng = np.random.RandomState(100 + seed)
    cam = S.intr_value
    ch = np.asarray(clone_handles, dtype=np.int32)
    fo, mo, mc, uvn, p0, cp0 = [0], [0], [], [], [], []
    Rc = jpl.quat_2_Rot(S.calib_value[:4])
    for k, pid in enumerate(S.plane_ids):
        if pid in drop_planes:
            continue
        idx = np.nonzero(S.planeid == pid)[0]
        for j, f in enumerate(idx):
            a, b = S.meas_offset[f], S.meas_offset[f + 1]
            if rng.rand() < slam_share and j > 0:
                a = b  # no measurements: a SLAM feature
            for q in range(a, b):
                mc.append(ch[S.meas_clone_idx[q]])
                if consistent:
                    v = S.clones[S.meas_clone_idx[q]][1]
                    pc = Rc @ (jpl.quat_2_Rot(v[:4]) @ (S.pf_true[f] - v[4:7])) + S.calib_value[4:7]
                    uvn.append(pc[:2] / pc[2] + px_noise / cam[0] * rng.randn(2))
                else:
                    uvn.append(vio_sim.undistort(cam, S.uv[q].astype(np.float64)))
            mo.append(mo[-1] + (b - a))
            p0.append(S.pf_true[f] + noise * rng.randn(3))  # triangulation error of a few centimetres
        fo.append(fo[-1] + len(idx))
        cp0.append(np.asarray(S.plane_cp)[k] * (1.0 + 0.01 * rng.randn()) + 0.004 * rng.randn(3))
    return dict(feat_offset=np.array(fo, dtype=np.int32), meas_offset=np.array(mo, dtype=np.int32), meas_clone=np.array(mc, dtype=np.int32),
                uv_norm=np.array(uvn, dtype=np.float32).reshape(-1, 2), p_FinG=np.array(p0).reshape(-1, 3), cp_inG=np.array(cp0).reshape(-1, 3))
