"""Deterministic synthetic clone-window scenarios for the MSCKF / point-on-plane update.

The reference ships no data that exercises N=256/512 states (its simulator caps at 11 clones); BASELINE.json's configs are
synthetic.  This generator follows the data model of the reference's simulator (`sim/Simulator.cpp`: EuRoC radtan camera,
ids offset by 4*max_aruco+1, plane id 0 = "no plane", on-plane points = ray/plane intersections, CP = n*d) and produces:
truth, an estimate drawn from a *valid correlated* prior covariance P0 (random-walk clone chain, reference priors for the
calibration blocks, `State.cpp:82-100`), first-estimate values, and a flat feature batch in the C-ABI's SoA layout.

Pure numpy, product-side utility: used by bench.py, __graft_entry__.smoke() and the tests (which feed the same scenario to
the CUDA path and to the oracle).
"""
import os
import numpy as np
from . import jpl

EUROC_CAM = np.array([458.654, 457.296, 367.215, 248.375, -0.28340811, 0.07395907, 0.00019359, 1.76187114e-05])
EUROC_WH = (752, 480)

CONFIGS = {
    # BASELINE.json configs[0] stand-in: shipped euroc_mav sizes (config/euroc_mav/estimator_config.yaml:12-19)
    "cfg1_euroc_n96": dict(n_clones=11, calib_pose=1, calib_intr=1, calib_dt=1, n_planes=0, n_slam=0, F=20, m_min=3, m_max=11,
                           plane_frac=0.0, dtheta=0.03),
    # configs[1]: N=256, 200 point features
    "cfg2_n256_f200": dict(n_clones=39, calib_pose=1, calib_intr=0, calib_dt=1, n_planes=0, n_slam=0, F=200, m_min=6, m_max=20,
                           plane_frac=0.0, dtheta=0.02),
    # configs[2]: N=512, 600 features + 8 in-state planes (+1 SLAM point to reach 512)
    "cfg3_n512_f600_p8": dict(n_clones=76, calib_pose=1, calib_intr=1, calib_dt=0, n_planes=8, n_slam=1, F=600, m_min=20, m_max=20,
                              plane_frac=0.5, dtheta=0.02),
    # configs[4]: N=488 (cfg3 without planes/SLAM), 4000 features to be sharded
    "cfg5_n512_f4000": dict(n_clones=76, calib_pose=1, calib_intr=1, calib_dt=0, n_planes=0, n_slam=0, F=4000, m_min=20, m_max=20,
                            plane_frac=0.0, dtheta=0.02),
    # small cases for fast parity tests
    "tiny_points": dict(n_clones=8, calib_pose=1, calib_intr=1, calib_dt=1, n_planes=0, n_slam=0, F=12, m_min=3, m_max=8,
                        plane_frac=0.0, dtheta=0.04),
    "tiny_planes": dict(n_clones=10, calib_pose=1, calib_intr=1, calib_dt=0, n_planes=2, n_slam=1, F=40, m_min=4, m_max=10,
                        plane_frac=0.6, dtheta=0.04),
    "small_planes": dict(n_clones=24, calib_pose=1, calib_intr=1, calib_dt=0, n_planes=4, n_slam=1, F=160, m_min=6, m_max=12,
                         plane_frac=0.5, dtheta=0.03),
}

_PLANES = [  # (normal, d) with n^T p = d, d > 0; CP = n * d   (room walls + ceiling/floor + interior walls)
    ((1.0, 0.0, 0.0), 5.0),
    ((0.0, 1.0, 0.0), 5.0),
    ((0.0, 0.0, 1.0), 3.0),
    ((0.0, 0.0, -1.0), 0.5),
    ((0.70710678118654757, 0.70710678118654757, 0.0), 4.9),
    ((-0.6, 0.8, 0.0), 4.6),
    ((1.0, 0.0, 0.0), 4.2),
    ((0.0, 1.0, 0.0), 4.4),
]


def chi2_table():
    path = os.path.join(os.path.dirname(__file__), "data", "chi2_095.txt")
    return np.loadtxt(path)


class Scenario(object):
    pass


def _project(cam, R_GtoI, p_I, R_ItoC, p_IinC, pf):
    pc = R_ItoC @ (R_GtoI @ (pf - p_I)) + p_IinC
    if pc[2] < 0.3:
        return None
    u, v = jpl.radtan_distort(cam, pc[0] / pc[2], pc[1] / pc[2])
    return u, v, pc[2]


def make_scenario(name="cfg3_n512_f600_p8", seed=0, sigma_px=1.0, **override):
    cfg = dict(CONFIGS[name])
    cfg.update(override)
    rng = np.random.RandomState(1234567 + 7919 * seed)
    S = Scenario()
    S.name, S.cfg, S.seed, S.sigma_px = name, cfg, seed, sigma_px
    C = cfg["n_clones"]
    S.options = dict(do_fej=1, imu_avg=0, use_rk4_integration=1, do_calib_camera_pose=cfg["calib_pose"],
                     do_calib_camera_intrinsics=cfg["calib_intr"], do_calib_camera_timeoffset=cfg["calib_dt"], max_clone_size=C,
                     max_aruco_features=1024, sigma_constraint=0.01, const_init_multi=1.0, const_init_chi2=1.0,
                     sigma_plane_merge=0.01, plane_merge_chi2=0.75, plane_merge_deg_max=1.0)
    W, H = EUROC_WH
    cam_true = EUROC_CAM.copy()
    R_ItoC_true = jpl.exp_so3(np.array([0.02, -0.01, 0.015]))
    p_IinC_true = np.array([0.02, -0.04, 0.01])

    # ---- true trajectory: camera on a circle looking outward -------------------------------------------------------
    th0, dth, rad = 0.35, cfg["dtheta"], 1.5
    R_true, p_true, ts = [], [], []
    for k in range(C):
        th = th0 + k * dth
        zc = np.array([np.cos(th), np.sin(th), 0.0])
        yc = np.array([0.0, 0.0, -1.0])
        xc = np.cross(yc, zc)
        R_GtoC = np.vstack([xc, yc, zc])
        R_GtoC = jpl.exp_so3(np.array([0.03 * np.sin(0.7 * k), 0.02 * np.cos(0.5 * k), 0.02 * np.sin(0.3 * k)])) @ R_GtoC
        R_GtoI = R_ItoC_true.T @ R_GtoC
        p = np.array([rad * np.cos(th), rad * np.sin(th), 1.5 + 0.1 * np.sin(0.4 * k)])
        R_true.append(R_GtoI)
        p_true.append(p)
        ts.append(100.0 + 0.05 * k)
    planes_true = []
    for i in range(cfg["n_planes"]):
        n, d = _PLANES[i % len(_PLANES)]
        n = np.array(n)
        n = n / np.linalg.norm(n)
        planes_true.append((i + 1, n * (d + 0.15 * (i // len(_PLANES)))))

    # ---- state layout + prior covariance P0 ----------------------------------------------------------------------
    layout = [("imu", 15)]
    if cfg["calib_dt"]:
        layout.append(("dt", 1))
    if cfg["calib_pose"]:
        layout.append(("calib", 6))
    if cfg["calib_intr"]:
        layout.append(("intr", 8))
    for k in range(C):
        layout.append(("clone%d" % k, 6))
    for pid, _ in planes_true:
        layout.append(("plane%d" % pid, 3))
    for i in range(cfg["n_slam"]):
        layout.append(("slam%d" % i, 3))
    off, ids = 0, {}
    for nme, sz in layout:
        ids[nme] = off
        off += sz
    N = off
    P0 = np.zeros((N, N))
    # kept small enough that second-order (linearisation) terms stay far below the pixel noise, so the 95% gates behave
    Pinit = np.diag([3e-3 ** 2] * 3 + [1e-2 ** 2] * 3)
    Qw = np.diag([4e-4 ** 2] * 3 + [1.5e-3 ** 2] * 3)
    for a in range(C):
        for b in range(C):
            blk = Pinit + (min(a, b) + 1) * Qw
            P0[ids["clone%d" % a]:ids["clone%d" % a] + 6, ids["clone%d" % b]:ids["clone%d" % b] + 6] = blk
    last = ids["clone%d" % (C - 1)]
    # IMU pose = newest clone (stochastic cloning just happened); v, bg, ba with own variances, mildly correlated to the pose
    P0[0:6, :] = P0[last:last + 6, :]
    P0[:, 0:6] = P0[:, last:last + 6]
    P0[0:6, 0:6] = P0[last:last + 6, last:last + 6]
    rest = np.diag([3e-2 ** 2] * 3 + [1e-3 ** 2] * 3 + [1e-2 ** 2] * 3)
    P0[6:15, 6:15] = rest
    G = 0.15 * rng.randn(9, 6)
    cross = np.sqrt(rest) @ G @ np.linalg.cholesky(P0[0:6, 0:6]).T
    P0[6:15, 0:6] = cross
    P0[0:6, 6:15] = cross.T
    P0[6:15, last:last + 6] = cross
    P0[last:last + 6, 6:15] = cross.T
    if cfg["calib_dt"]:
        P0[ids["dt"], ids["dt"]] = 0.01 ** 2
    if cfg["calib_pose"]:
        b = ids["calib"]
        P0[b:b + 6, b:b + 6] = np.diag([0.005 ** 2] * 3 + [0.01 ** 2] * 3)
    if cfg["calib_intr"]:
        b = ids["intr"]
        P0[b:b + 8, b:b + 8] = np.diag([1.0 ** 2] * 4 + [0.005 ** 2] * 4)
    for pid, _ in planes_true:
        b = ids["plane%d" % pid]
        P0[b:b + 3, b:b + 3] = 0.03 ** 2 * np.eye(3)
    for i in range(cfg["n_slam"]):
        b = ids["slam%d" % i]
        P0[b:b + 3, b:b + 3] = 0.05 ** 2 * np.eye(3)
    P0 = 0.5 * (P0 + P0.T)
    # the pose rows of imu and newest clone are identical => singular; add a tiny propagation noise to the IMU block
    P0[0:15, 0:15] += np.diag([1e-8] * 15)
    w = np.linalg.eigvalsh(P0)
    if w[0] <= 0:
        P0 += (abs(w[0]) + 1e-12) * np.eye(N)
    L0 = np.linalg.cholesky(P0)
    err = L0 @ rng.randn(N)

    # ---- estimates = truth (+) error; first-estimates = estimate (+) small drift ------------------------------------
    def pose_val(R, p):
        return np.concatenate([jpl.rot_2_quat(R), p])

    def perturb_pose(val, d):
        q = jpl.quat_left_update(val[:4], d[:3])
        return np.concatenate([q, val[4:7] + d[3:6]])

    S.timestamp = ts[-1]
    S.clones = []
    for k in range(C):
        tv = pose_val(R_true[k], p_true[k])
        ev = perturb_pose(tv, err[ids["clone%d" % k]:ids["clone%d" % k] + 6])
        fv = perturb_pose(ev, np.concatenate([8e-4 * rng.randn(3), 2e-3 * rng.randn(3)]))
        S.clones.append((ts[k], ev, fv))
    imu_true = np.concatenate([pose_val(R_true[-1], p_true[-1]), np.array([0.1, 0.05, 0.0]), np.zeros(3), np.zeros(3)])
    e = err[0:15]
    imu_est = np.concatenate([perturb_pose(imu_true[:7], e[:6]), imu_true[7:16] + e[6:15]])
    S.imu_value, S.imu_fej = imu_est, imu_est.copy()
    S.dt_value = np.array([0.0 + (err[ids["dt"]] if cfg["calib_dt"] else 0.0)])
    calib_true = pose_val(R_ItoC_true, p_IinC_true)
    S.calib_value = perturb_pose(calib_true, err[ids["calib"]:ids["calib"] + 6]) if cfg["calib_pose"] else calib_true
    S.intr_value = cam_true + (err[ids["intr"]:ids["intr"] + 8] if cfg["calib_intr"] else 0.0)
    S.planes = []
    for pid, cp in planes_true:
        ev = cp + err[ids["plane%d" % pid]:ids["plane%d" % pid] + 3]
        S.planes.append((pid, ev, ev + 1e-3 * rng.randn(3)))
    S.slam = []
    for i in range(cfg["n_slam"]):
        pt = p_true[C // 2] + 3.0 * np.array([np.cos(th0 + dth * C / 2), np.sin(th0 + dth * C / 2), 0.1])
        ev = pt + err[ids["slam%d" % i]:ids["slam%d" % i] + 3]
        S.slam.append((5000000 + i, ev, ev + 1e-3 * rng.randn(3)))
    S.P0, S.N, S.ids, S.layout = P0, N, ids, layout

    # ---- features ----------------------------------------------------------------------------------------------------
    F = cfg["F"]
    meas_offset = [0]
    meas_clone_idx, uv, pf_true_l, planeid_l = [], [], [], []
    nplane_feats = 0
    Rt = np.stack(R_true)
    pt_ = np.stack(p_true)
    fx, fy, cx, cy = cam_true[:4]
    for i in range(F):
        on_plane = cfg["n_planes"] > 0 and rng.rand() < cfg["plane_frac"]
        pid = 0
        if on_plane:
            pid = 1 + (nplane_feats % cfg["n_planes"])
        ok = False
        for _try in range(400):
            m = rng.randint(cfg["m_min"], min(cfg["m_max"], C) + 1)
            e_ = rng.randint(m - 1, C)
            win = list(range(e_ - m + 1, e_ + 1))
            cref = win[len(win) // 2]
            u0 = rng.uniform(60, W - 60)
            v0 = rng.uniform(60, H - 60)
            ray_c = np.array([(u0 - cx) / fx, (v0 - cy) / fy, 1.0])
            R_GtoC = R_ItoC_true @ Rt[cref]
            pc_G = pt_[cref] - Rt[cref].T @ (R_ItoC_true.T @ p_IinC_true)
            ray_G = R_GtoC.T @ ray_c
            if pid:
                cp = planes_true[pid - 1][1]
                d = np.linalg.norm(cp)
                n = cp / d
                den = n @ ray_G
                if abs(den) < 1e-3:
                    continue
                t = (d - n @ pc_G) / den
                if t < 1.5 or t > 12.0:
                    continue
            else:
                t = rng.uniform(2.0, 5.0)
            pf = pc_G + t * ray_G
            pix = []
            good = True
            for k in win:
                pr = _project(cam_true, Rt[k], pt_[k], R_ItoC_true, p_IinC_true, pf)
                if pr is None or not (8 < pr[0] < W - 8 and 8 < pr[1] < H - 8):
                    good = False
                    break
                pix.append((pr[0], pr[1]))
            if not good:
                continue
            ok = True
            break
        if not ok:
            raise RuntimeError("could not place feature %d" % i)
        if pid:
            nplane_feats += 1
        for k, (pu, pv) in zip(win, pix):
            meas_clone_idx.append(k)
            uv.append((pu + sigma_px * rng.randn(), pv + sigma_px * rng.randn()))
        meas_offset.append(len(meas_clone_idx))
        pf_true_l.append(pf)
        planeid_l.append(pid)
    S.F = F
    S.meas_offset = np.array(meas_offset, dtype=np.int32)
    S.meas_clone_idx = np.array(meas_clone_idx, dtype=np.int32)  # index into S.clones; backends map it to handles
    S.uv = np.array(uv, dtype=np.float64).astype(np.float32).reshape(-1, 2)
    S.pf_true = np.array(pf_true_l).reshape(-1, 3)
    S.planeid = np.array(planeid_l, dtype=np.int64)
    S.featid = np.arange(F, dtype=np.int64) + 4 * 1024 + 1
    noisy = S.pf_true + 0.02 * rng.randn(F, 3)
    S.p_FinG_original = noisy.copy()
    refined = noisy.copy()
    est_planes = {pid: ev for pid, ev, _ in S.planes}
    for i in range(F):
        if S.planeid[i]:
            cp = est_planes[int(S.planeid[i])]
            d = np.linalg.norm(cp)
            n = cp / d
            refined[i] = noisy[i] - 0.9 * n * (n @ noisy[i] - d)  # stand-in for the plane refinement (UpdaterMSCKF.cpp:278-280)
    S.p_FinG = refined
    S.plane_ids = np.array([pid for pid, _, _ in S.planes], dtype=np.int64)
    S.plane_cp = np.array([ev for _, ev, _ in S.planes], dtype=np.float64).reshape(-1, 3)
    return S


def load_scenario_into(backend, S):
    """Feed a scenario into an object exposing the C-ABI's method names (ov_plane_b200.api.Context or the oracle mirror in
    tests/).  Returns the list of clone handles (index = clone number)."""
    backend.var_set(backend.handle_imu(), S.imu_value, S.imu_fej)
    backend.var_set(backend.handle_dt(), S.dt_value, S.dt_value)
    backend.var_set(backend.handle_calib(), S.calib_value, S.calib_value)
    backend.var_set(backend.handle_intrinsics(), S.intr_value, S.intr_value)
    clone_handles = []
    for (t, v, f) in S.clones:
        clone_handles.append(backend.add_clone_raw(t, v, f))
    for (pid, v, f) in S.planes:
        backend.add_plane_raw(pid, v, f)
    for (fid, v, f) in S.slam:
        backend.add_slam_raw(fid, v, f)
    backend.set_timestamp(S.timestamp)
    assert backend.cov_rows() == S.N, (backend.cov_rows(), S.N)
    backend.cov_upload(S.P0)
    return clone_handles


def feature_batch(S, clone_handles, sel=None):
    """SoA batch (dict of contiguous numpy arrays) in the layout of `ovp_feature_batch`; sel = optional feature subset."""
    idx = np.arange(S.F) if sel is None else np.asarray(sel)
    offs = [0]
    mc, uv = [], []
    ch = np.asarray(clone_handles, dtype=np.int32)
    for i in idx:
        a, b = S.meas_offset[i], S.meas_offset[i + 1]
        mc.append(ch[S.meas_clone_idx[a:b]])
        uv.append(S.uv[a:b])
        offs.append(offs[-1] + (b - a))
    return dict(
        F=len(idx), meas_offset=np.array(offs, dtype=np.int32),
        meas_clone=np.ascontiguousarray(np.concatenate(mc) if mc else np.zeros(0), dtype=np.int32),
        uv=np.ascontiguousarray(np.concatenate(uv) if uv else np.zeros((0, 2)), dtype=np.float32),
        p_FinG=np.ascontiguousarray(S.p_FinG[idx], dtype=np.float64),
        p_FinG_original=np.ascontiguousarray(S.p_FinG_original[idx], dtype=np.float64),
        featid=np.ascontiguousarray(S.featid[idx], dtype=np.int64), planeid=np.ascontiguousarray(S.planeid[idx], dtype=np.int64),
        plane_ids=np.ascontiguousarray(S.plane_ids, dtype=np.int64), plane_cp=np.ascontiguousarray(S.plane_cp, dtype=np.float64))


def drop_planes_from_state(S):
    """Variant of a scenario whose planes are NOT state variables (fresh planes: MSCKF update with the plane projected away,
    UpdaterMSCKF.cpp:601-604, and plane initialisation, UpdaterPlane.cpp:297-481).  Returns the removed (id, cp, cp_fej) list."""
    planes = S.planes
    drop = set()
    for pid, _, _ in planes:
        b = S.ids["plane%d" % pid]
        drop.update(range(b, b + 3))
    keep = [i for i in range(S.N) if i not in drop]
    S.P0 = np.ascontiguousarray(S.P0[np.ix_(keep, keep)])
    S.N = len(keep)
    S.planes = []
    return planes
