"""ROS-free VioManager loop + simulator for BASELINE config 4 ("room" trajectory, 30-clone window, drop-in under VioManager).

Two parts, both driving ONLY the C ABI surface (any backend with the method names of `ov_plane_b200.api.Context`: the CUDA library,
or the CPU oracle mirror in tests/ for parity runs):

* `RoomSimulator` — data model of the reference's simulator (`ov_plane/src/sim/Simulator.cpp:143-155, 436-439, 580-605, 645-707`):
  a smooth SE(3) trajectory inside a 6-plane room, IMU samples at 400 Hz with the reference's noise model
  (`utils/NoiseManager.h:41-63`: white noise + bias random walks), a 10 Hz mono radtan camera, features that are either ray/plane
  intersections (plane id 1..6) or free points (plane id 0), pixel noise sigma_px.  The reference drives a B-spline through a recorded
  trajectory file (`data/udel_room_05.txt`, ov_core `BsplineSE3`, neither available on the GPU box); the trajectory here is an
  analytic curve of the same scale (2 m loop, 1.5 m height, ~0.4 m/s) differentiated numerically.
* `VioLoop` — the call order of `VioManager::do_feature_propagate_update` (`ov_plane/src/core/VioManager.cpp:330-986`):
  propagate_and_clone (:347) -> feature selection lost / marg / max-track (:373-506) -> feat2plane from the simulator (:221-230),
  plane bookkeeping (:516-534) -> init_vio_plane (:585-588) -> UpdaterMSCKF::update (:670) -> marginalize_old_clone (:872), with the
  per-stage timing columns of the reference's CSV (:110-118, 911-928).  Triangulation (ov_core FeatureInitializer: linear + 5
  Gauss-Newton steps) is the host-side stand-in of the stage upstream of the path, the plane fit is supplied by the simulator
  upstream of the path.
"""
import time

import numpy as np

from . import jpl
from .synth import EUROC_CAM, EUROC_WH, _project

GRAVITY = np.array([0.0, 0.0, 9.81])
ROOM_PLANES = [  # (plane id, normal, d) with n^T p = d > 0; CP = n * d  (Simulator.cpp:645-707: a closed room)
    (1, (1.0, 0.0, 0.0), 5.0), (2, (-1.0, 0.0, 0.0), 5.0), (3, (0.0, 1.0, 0.0), 5.0), (4, (0.0, -1.0, 0.0), 5.0),
    (5, (0.0, 0.0, 1.0), 3.0), (6, (0.0, 0.0, -1.0), 0.5)]


class RoomSimulator(object):
    def __init__(self, seed=0, sigma_px=1.0, cam_hz=10.0, imu_hz=400.0, n_feats=60, plane_frac=0.6,
                 sigma_w=1.6968e-04, sigma_wb=1.9393e-05, sigma_a=2.0e-3, sigma_ab=3.0e-3):
        self.rng = np.random.RandomState(9000 + seed)
        self.sigma_px, self.cam_dt, self.imu_dt = sigma_px, 1.0 / cam_hz, 1.0 / imu_hz
        self.n_feats, self.plane_frac = n_feats, plane_frac
        self.noise = (sigma_w, sigma_wb, sigma_a, sigma_ab)
        self.cam = EUROC_CAM.copy()
        self.R_ItoC = jpl.exp_so3(np.array([0.02, -0.01, 0.015]))
        self.p_IinC = np.array([0.02, -0.04, 0.01])
        self.bg = np.zeros(3)
        self.ba = np.zeros(3)
        self.t_imu = None
        self.next_id = 4 * 1024 + 1  # ids above 4 * max_aruco (VioManager.cpp:226-228)
        self.active = {}  # featid -> (p_FinG, planeid)

    # ---- trajectory: camera on a 2 m loop looking outward, gentle roll / pitch / height oscillation ----
    def pose(self, t):
        th = 0.35 + 0.2 * t
        zc = np.array([np.cos(th), np.sin(th), 0.0])
        yc = np.array([0.0, 0.0, -1.0])
        xc = np.cross(yc, zc)
        R_GtoC = np.vstack([xc, yc, zc])
        R_GtoC = jpl.exp_so3(np.array([0.06 * np.sin(0.9 * t), 0.05 * np.cos(0.7 * t), 0.04 * np.sin(0.5 * t)])) @ R_GtoC
        R_GtoI = self.R_ItoC.T @ R_GtoC
        p = np.array([2.0 * np.cos(th), 2.0 * np.sin(th), 1.5 + 0.15 * np.sin(0.8 * t)])
        return R_GtoI, p

    def kinematics(self, t):
        """(R_GtoI, p, v, body angular velocity, specific force) by central differences of the analytic pose"""
        h = 1e-4
        R0, p0 = self.pose(t)
        Rp, pp = self.pose(t + h)
        Rm, pm = self.pose(t - h)
        v = (pp - pm) / (2 * h)
        acc = (pp - 2 * p0 + pm) / (h * h)
        # JPL: d/dt R_GtoI = -skew(w) R_GtoI  =>  skew(w) = -(dR/dt) R^T
        dR = (Rp - Rm) / (2 * h)
        W = -dR @ R0.T
        w = np.array([W[2, 1] - W[1, 2], W[0, 2] - W[2, 0], W[1, 0] - W[0, 1]]) * 0.5
        return R0, p0, v, w, R0 @ (acc + GRAVITY)

    def imu_until(self, t_end):
        """IMU samples (t, wm, am) up to and one past t_end (the propagator interpolates the last one, Propagator.cpp:226-341)"""
        sw, swb, sa, sab = self.noise
        out = []
        if self.t_imu is None:
            self.t_imu = 0.0
        while self.t_imu <= t_end + 2 * self.imu_dt:
            _, _, _, w, a = self.kinematics(self.t_imu)
            dt = self.imu_dt
            wm = w + self.bg + sw / np.sqrt(dt) * self.rng.randn(3)
            am = a + self.ba + sa / np.sqrt(dt) * self.rng.randn(3)
            self.bg = self.bg + swb * np.sqrt(dt) * self.rng.randn(3)
            self.ba = self.ba + sab * np.sqrt(dt) * self.rng.randn(3)
            out.append((self.t_imu, wm, am))
            self.t_imu += dt
        return out

    def _new_feature(self, R_GtoI, p):
        W, H = EUROC_WH
        fx, fy, cx, cy = self.cam[:4]
        for _ in range(200):
            u0, v0 = self.rng.uniform(40, W - 40), self.rng.uniform(40, H - 40)
            ray_c = np.array([(u0 - cx) / fx, (v0 - cy) / fy, 1.0])
            R_GtoC = self.R_ItoC @ R_GtoI
            pc_G = p - R_GtoI.T @ (self.R_ItoC.T @ self.p_IinC)
            ray_G = R_GtoC.T @ ray_c
            if self.rng.rand() < self.plane_frac:
                best = None
                for pid, n, d in ROOM_PLANES:  # first wall hit by the ray (Simulator.cpp:580-605)
                    n = np.asarray(n)
                    den = n @ ray_G
                    if den < 1e-3:
                        continue
                    tt = (d - n @ pc_G) / den
                    if tt > 0.5 and (best is None or tt < best[0]):
                        best = (tt, pid)
                if best is None or best[0] > 12.0:
                    continue
                return pc_G + best[0] * ray_G, best[1]
            return pc_G + self.rng.uniform(2.0, 4.5) * ray_G, 0
        raise RuntimeError("could not place a feature")

    def camera_frame(self, t):
        """{featid: (uv float32 (2,), planeid)} of the features visible at time t; the feature set is topped up to n_feats"""
        R, p, _, _, _ = self.kinematics(t)
        W, H = EUROC_WH
        obs = {}
        for fid in list(self.active):
            pf, pid = self.active[fid]
            pr = _project(self.cam, R, p, self.R_ItoC, self.p_IinC, pf)
            if pr is None or not (5 < pr[0] < W - 5 and 5 < pr[1] < H - 5):
                del self.active[fid]
                continue
            obs[fid] = (np.array([pr[0] + self.sigma_px * self.rng.randn(), pr[1] + self.sigma_px * self.rng.randn()], dtype=np.float32), pid)
        while len(obs) < self.n_feats:
            pf, pid = self._new_feature(R, p)
            pr = _project(self.cam, R, p, self.R_ItoC, self.p_IinC, pf)
            if pr is None:
                continue
            fid = self.next_id
            self.next_id += 1
            self.active[fid] = (pf, pid)
            obs[fid] = (np.array([pr[0] + self.sigma_px * self.rng.randn(), pr[1] + self.sigma_px * self.rng.randn()], dtype=np.float32), pid)
        return obs


def state_options(max_clones=30, calib=True):
    return dict(do_fej=1, imu_avg=0, use_rk4_integration=1, do_calib_camera_pose=int(calib), do_calib_camera_intrinsics=int(calib),
                do_calib_camera_timeoffset=0, max_clone_size=max_clones, max_aruco_features=1024, sigma_constraint=0.01, const_init_multi=1.0,
                const_init_chi2=1.0, sigma_plane_merge=0.01, plane_merge_chi2=0.75, plane_merge_deg_max=1.0)


def triangulate(poses, uvn):
    """ov_core FeatureInitializer::single_triangulation + single_gaussnewton (UpdaterMSCKF.cpp:142-194), host side: linear
    least squares on the bearing cross products, then 5 Gauss-Newton steps on the normalised reprojection error.
    poses: list of (R_GtoC, p_CinG); uvn: normalised coordinates."""
    A, b = np.zeros((3, 3)), np.zeros(3)
    for (R, pc), z in zip(poses, uvn):
        bear = R.T @ np.array([z[0], z[1], 1.0])
        bear /= np.linalg.norm(bear)
        Bp = np.eye(3) - np.outer(bear, bear)
        A += Bp
        b += Bp @ pc
    pf = np.linalg.solve(A, b)
    for _ in range(5):
        J, r = [], []
        for (R, pc), z in zip(poses, uvn):
            q = R @ (pf - pc)
            if q[2] < 0.1:
                return None
            r.append(np.array([q[0] / q[2] - z[0], q[1] / q[2] - z[1]]))
            J.append(np.array([[1 / q[2], 0, -q[0] / q[2] ** 2], [0, 1 / q[2], -q[1] / q[2] ** 2]]) @ R)
        J, r = np.vstack(J), np.concatenate(r)
        pf = pf - np.linalg.solve(J.T @ J + 1e-9 * np.eye(3), J.T @ r)
    return pf


def undistort(cam, uv):
    """radtan inverse by fixed-point iteration (ov_core CamRadtan::undistort uses OpenCV's; 8 iterations reach 1e-12 here)"""
    fx, fy, cx, cy, k1, k2, p1, p2 = cam
    x0, y0 = (uv[0] - cx) / fx, (uv[1] - cy) / fy
    x, y = x0, y0
    for _ in range(8):
        r2 = x * x + y * y
        rad = 1 + k1 * r2 + k2 * r2 * r2
        dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
        dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
        x, y = (x0 - dx) / rad, (y0 - dy) / rad
    return np.array([x, y])


class VioLoop(object):
    """VioManager::do_feature_propagate_update against one backend.  `gate_ctx`: optional context-manager factory wrapped around the
    updater calls (a parity run wraps the CPU checker's calls so that its plane gates use the well-defined chi2, see tests/)."""

    def __init__(self, backend, sim, max_clones=30, min_clones_for_update=5, plane_init_min_feat=8, chi2_mult=1.0, gate_ctx=None,
                 use_planes=True, fit_planes=False, sigma_c=0.01):
        self.be, self.sim = backend, sim
        self.max_clones, self.min_clones = max_clones, min_clones_for_update
        self.plane_init_min_feat, self.chi2_mult, self.use_planes = plane_init_min_feat, chi2_mult, use_planes
        self.gate_ctx = gate_ctx
        # fit_planes: plane estimates and refined feature positions come from PlaneFitting::plane_fitting / optimize_plane through the backend
        # (UpdaterPlane.cpp:224-270, UpdaterMSCKF.cpp:262-360) instead of the simulator's stand-in
        self.fit_planes, self.sigma_c = fit_planes, sigma_c
        self.fit_stats = dict(ransac_ok=0, ransac_fail=0, refine_ok=0, refine_fail=0)
        self.tracks = {}   # featid -> list of (t, uv)
        self.planeof = {}  # featid -> planeid
        self.clone_times = []
        self.clone_handles = {}  # clone time -> handle
        self.plane_fit = {}      # plane id -> closest-point estimate handed to init_vio_plane
        self.frames = []   # per-frame records
        sw, swb, sa, sab = sim.noise
        backend.propagator_set_noise(sw, swb, sa, sab, 9.81)

    def initialize_with_gt(self, t0):
        """VioManager::initialize_with_gt (run_simulation.cpp:104-120): IMU state = truth, small diagonal prior"""
        be, sim = self.be, self.sim
        R, p, v, _, _ = sim.kinematics(t0)
        val = np.concatenate([jpl.rot_2_quat(R), p, v, np.zeros(3), np.zeros(3)])
        be.var_set(be.handle_imu(), val, val)
        calib = np.concatenate([jpl.rot_2_quat(sim.R_ItoC), sim.p_IinC])
        be.var_set(be.handle_calib(), calib, calib)
        be.var_set(be.handle_intrinsics(), sim.cam, sim.cam)
        be.set_timestamp(t0)
        N = be.cov_rows()
        P = np.zeros((N, N))
        d = np.array([1e-3] * 3 + [1e-3] * 3 + [1e-2] * 3 + [1e-3] * 3 + [1e-2] * 3) ** 2
        P[:15, :15] = np.diag(d)
        k = 15
        if N >= 15 + 6:
            P[k:k + 6, k:k + 6] = np.diag([0.003 ** 2] * 3 + [0.005 ** 2] * 3)
            k += 6
        if N >= k + 8:
            P[k:k + 8, k:k + 8] = np.diag([0.5 ** 2] * 4 + [0.002 ** 2] * 4)
        be.cov_upload(P)
        self.t_last_imu = -1.0

    def _feed_imu(self, t):
        for (ti, wm, am) in self.sim.imu_until(t):
            self.be.feed_imu(ti, wm, am)

    def step(self, t, obs):
        """one camera frame: obs = simulator output {featid: (uv, planeid)}"""
        be = self.be
        rec = {"t": t}
        T0 = time.perf_counter()
        for fid, (uv, pid) in obs.items():
            self.tracks.setdefault(fid, []).append((t, uv))
            self.planeof[fid] = pid
        # ---- 1. propagation + clone (VioManager.cpp:347-349) ----
        self.clone_handles[t] = be.propagate_and_clone(t)[0]
        self.clone_times.append(t)
        T1 = time.perf_counter()
        rec["n_clones"] = len(self.clone_times)
        if len(self.clone_times) < self.min_clones:  # :353-361
            rec.update(propagation=T1 - T0, front_end=0.0, plane_init=0.0, msckf=0.0, marg=0.0, n_msckf=0, n_planes=self._nplanes())
            self._record(rec)
            return rec
        # ---- 2. feature selection (:373-506): lost tracks, tracks touching the clone about to be marginalised, max-length tracks ----
        marg_t = self.clone_times[0] if len(self.clone_times) > self.max_clones else None
        sel = []
        for fid, tr in self.tracks.items():
            lost = tr[-1][0] < t
            hits_marg = marg_t is not None and any(abs(tt - marg_t) < 1e-9 for tt, _ in tr)
            if lost or hits_marg or len(tr) > self.max_clones:
                sel.append(fid)
        sel.sort(key=lambda f: len(self.tracks[f]))  # ascending track length (:608-623)
        batch, used = self._build_batch(sel)
        T1b = time.perf_counter()  # feature selection, host triangulation and (fit_planes) the plane fits are the front end: the "tracking" column
        T2 = T1b
        n_used = 0
        if batch is not None:
            ctxm = self.gate_ctx() if self.gate_ctx else _null()
            with ctxm:
                # ---- 4. plane initialisation (:585-588), then 5. the MSCKF update (:670) ----
                if self.use_planes and len(batch["plane_ids"]):
                    r0 = be.plane_init(batch, self.sim.sigma_px)
                    rec["planes_initialised"] = int((np.asarray(r0["plane_status"]) == 1).sum())
                T2 = time.perf_counter()
                out = be.msckf_update(batch, self.sim.sigma_px, self.chi2_mult)
            rec["feat_status"] = out["feat_status"].copy()
            rec["plane_status"] = out["plane_status"].copy()
            n_used = int((out["feat_status"] >= 1).sum())
        T3 = time.perf_counter()
        for fid in sel:  # every selected feature is consumed (to_delete) whether accepted or not
            self.tracks.pop(fid, None)
        # ---- 7. marginalisation (:872) + planes without active features (:516-534) ----
        if len(self.clone_times) > self.max_clones:
            be.marginalize_old_clone()
            old = self.clone_times.pop(0)
            self.clone_handles.pop(old, None)
            for fid in list(self.tracks):
                self.tracks[fid] = [(tt, uv) for tt, uv in self.tracks[fid] if tt > old + 1e-9]
                if not self.tracks[fid]:
                    del self.tracks[fid]
        if self.use_planes:
            f2p = {int(f): int(self.planeof[f]) for f in self.tracks if self.planeof.get(f, 0) > 0}
            be.merge_planes_and_marginalize(f2p, {})
        T4 = time.perf_counter()
        rec.update(propagation=T1 - T0, front_end=T1b - T1, plane_init=T2 - T1b, msckf=T3 - T2, marg=T4 - T3, n_msckf=len(sel), n_used=n_used,
                   n_planes=self._nplanes())
        self._record(rec)
        return rec

    def _nplanes(self):
        return sum(1 for pid, _, _ in ROOM_PLANES if self.be.plane_handle(pid) >= 0)

    def _build_batch(self, sel):
        be, sim = self.be, self.sim
        calib, _ = be.var_get(be.handle_calib())
        cam, _ = be.var_get(be.handle_intrinsics())
        R_ItoC, p_IinC = jpl.quat_2_Rot(calib[:4]), calib[4:7]
        hmap, cpose = {}, {}
        for tt in self.clone_times:
            h = self.clone_handles[tt]
            v, _ = be.var_get(h)
            Rg = jpl.quat_2_Rot(v[:4])
            hmap[tt] = h
            cpose[tt] = (R_ItoC @ Rg, v[4:7] - Rg.T @ (R_ItoC.T @ p_IinC))
        offs, mcl, uvs, pfs, fids, pids, anchors = [0], [], [], [], [], [], []
        for fid in sel:
            tr = [(tt, uv) for tt, uv in self.tracks[fid] if tt in hmap]
            if len(tr) < 3:
                continue
            tr = tr[-min(len(tr), 32):]
            pf = triangulate([cpose[tt] for tt, _ in tr], [undistort(cam, uv.astype(np.float64)) for _, uv in tr])
            if pf is None or not np.all(np.isfinite(pf)):
                continue
            depth = (cpose[tr[-1][0]][0] @ (pf - cpose[tr[-1][0]][1]))[2]
            if depth < 0.25 or depth > 40.0:  # FeatureInitializer max_dist / min_dist
                continue
            for tt, uv in tr:
                mcl.append(hmap[tt])
                uvs.append(uv)
            offs.append(len(mcl))
            anchors.append(cpose[tr[len(tr) // 2][0]][1])
            pfs.append(pf)
            fids.append(fid)
            pids.append(self.planeof.get(fid, 0) if self.use_planes else 0)
        if not fids:
            return None, []
        pfs = np.array(pfs)
        pids = np.array(pids, dtype=np.int64)
        if self.fit_planes:
            return self._finish_batch_with_plane_fitting(offs, mcl, uvs, pfs, fids, pids, cam)
        # plane estimates: in-state planes come from the state; new planes get a least-squares fit through their triangulated points
        plane_ids, plane_cp = [], []
        for pid in sorted(set(int(x) for x in pids if x > 0)):
            idx = np.nonzero(pids == pid)[0]
            if be.plane_handle(pid) >= 0:
                plane_ids.append(pid)
                plane_cp.append(be.var_get(be.plane_handle(pid))[0][:3])
                continue
            if len(idx) < self.plane_init_min_feat:
                pids[idx] = 0  # too few points to fit / initialise: treat as ordinary point features this frame
                continue
            # the RANSAC + Ceres plane fit (PlaneFitting.cpp:83-514) is upstream of the path: the simulator hands the loop an estimate
            # of that quality (truth + 2 cm), drawn once per plane so that both backends of a parity run see the same value
            if pid not in self.plane_fit:
                n, d = next((np.asarray(nn), dd) for pp, nn, dd in ROOM_PLANES if pp == pid)
                self.plane_fit[pid] = n * d + 0.02 * np.random.RandomState(77 + pid).randn(3)
            plane_ids.append(pid)
            plane_cp.append(self.plane_fit[pid])
        # refined positions (stand-in of optimize_plane with the plane fixed, PlaneFitting.cpp:197-514): slide the triangulated point
        # along the viewing ray of its middle observation until it lies on the plane estimate - the bearings stay satisfied
        pref = pfs.copy()
        for pid, cp in zip(plane_ids, plane_cp):
            d = np.linalg.norm(cp)
            n = cp / d
            for i in np.nonzero(pids == pid)[0]:
                pc = anchors[i]
                ray = pfs[i] - pc
                den = n @ ray
                if abs(den) > 1e-6:
                    tt = (d - n @ pc) / den
                    if 0.5 < tt < 2.0:
                        pref[i] = pc + tt * ray
        batch = dict(F=len(fids), meas_offset=np.array(offs, dtype=np.int32), meas_clone=np.array(mcl, dtype=np.int32),
                     uv=np.ascontiguousarray(np.array(uvs, dtype=np.float32).reshape(-1, 2)), p_FinG=np.ascontiguousarray(pref),
                     p_FinG_original=np.ascontiguousarray(pfs), featid=np.array(fids, dtype=np.int64), planeid=pids,
                     plane_ids=np.array(plane_ids, dtype=np.int64), plane_cp=np.ascontiguousarray(np.array(plane_cp, dtype=np.float64).reshape(-1, 3)))
        return batch, fids

    def _finish_batch_with_plane_fitting(self, offs, mcl, uvs, pfs, fids, pids, cam):
        """The reference's order of business for the planes of a frame: a plane that is NOT in the state gets a RANSAC hypothesis from its
        triangulated points (plane_fitting) and a joint refinement of plane + points (optimize_plane, free plane); a plane that IS in the state
        only refines its points against the state's estimate (fixed plane).  Whatever fails falls back to plain point features this frame."""
        be = self.be
        pids = pids.copy()
        offs = np.asarray(offs, dtype=np.int32)
        uvn = np.array([undistort(cam, np.asarray(uv, dtype=np.float64)) for uv in uvs], dtype=np.float32).reshape(-1, 2)
        groups = {int(pid): np.nonzero(pids == pid)[0] for pid in sorted(set(int(x) for x in pids if x > 0))}
        new = [pid for pid, idx in groups.items() if be.plane_handle(pid) < 0 and len(idx) >= self.plane_init_min_feat]
        for pid, idx in groups.items():
            if be.plane_handle(pid) < 0 and pid not in new:
                pids[idx] = 0
        cp_of = {}
        if new:
            fo = np.cumsum([0] + [len(groups[pid]) for pid in new]).astype(np.int32)
            st, ab, inl = be.plane_fitting(fo, np.vstack([pfs[groups[pid]] for pid in new]), self.plane_init_min_feat, 200.0)
            for k, pid in enumerate(new):
                idx = groups[pid]
                if not st[k]:
                    self.fit_stats["ransac_fail"] += 1
                    pids[idx] = 0
                    continue
                self.fit_stats["ransac_ok"] += 1
                keep = inl[fo[k]:fo[k + 1]] == 1
                pids[idx[~keep]] = 0
                groups[pid] = idx[keep]
                cp_of[pid] = -ab[k, :3] * ab[k, 3]
        cand = [pid for pid in groups if (pids[groups[pid]] == pid).any() and (pid in cp_of or be.plane_handle(pid) >= 0)]
        pref = pfs.copy()
        plane_ids, plane_cp = [], []
        if cand:
            fo, mo, mc, uv2, p0, cp0, fx = [0], [0], [], [], [], [], []
            for pid in cand:
                idx = groups[pid]
                for i in idx:
                    a, b = offs[i], offs[i + 1]
                    mc.extend(mcl[a:b])
                    uv2.append(uvn[a:b])
                    mo.append(mo[-1] + (b - a))
                    p0.append(pfs[i])
                fo.append(fo[-1] + len(idx))
                in_state = be.plane_handle(pid) >= 0
                cp0.append(be.var_get(be.plane_handle(pid))[0][:3] if in_state else cp_of[pid])
                fx.append(1 if in_state else 0)
            st, po, co, inl, _ = be.optimize_plane(np.array(fo, dtype=np.int32), np.array(mo, dtype=np.int32), np.array(mc, dtype=np.int32),
                                                   np.vstack(uv2), np.array(p0), np.array(cp0), np.array(fx, dtype=np.int32),
                                                   self.sim.sigma_px / cam[0], self.sigma_c)
            for k, pid in enumerate(cand):
                idx = groups[pid]
                if not st[k]:  # the reference skips the plane for this frame (`continue`, UpdaterMSCKF.cpp:279-280, 353-354)
                    self.fit_stats["refine_fail"] += 1
                    pids[idx] = 0
                    continue
                self.fit_stats["refine_ok"] += 1
                keep = inl[fo[k]:fo[k + 1]] == 1
                pref[idx[keep]] = po[fo[k]:fo[k + 1]][keep]
                pids[idx[~keep]] = 0
                plane_ids.append(pid)
                plane_cp.append(co[k] if not fx[k] else np.asarray(cp0[k]))
        batch = dict(F=len(fids), meas_offset=offs, meas_clone=np.array(mcl, dtype=np.int32),
                     uv=np.ascontiguousarray(np.array(uvs, dtype=np.float32).reshape(-1, 2)), p_FinG=np.ascontiguousarray(pref),
                     p_FinG_original=np.ascontiguousarray(pfs), featid=np.array(fids, dtype=np.int64), planeid=pids,
                     plane_ids=np.array(plane_ids, dtype=np.int64), plane_cp=np.ascontiguousarray(np.array(plane_cp, dtype=np.float64).reshape(-1, 3)))
        return batch, fids

    def _record(self, rec):
        be, sim = self.be, self.sim
        v, _ = be.var_get(be.handle_imu())
        R, p, vel, _, _ = sim.kinematics(rec["t"])
        Re = jpl.quat_2_Rot(v[:4])
        dR = Re @ R.T
        th = 0.5 * np.array([dR[1, 2] - dR[2, 1], dR[2, 0] - dR[0, 2], dR[0, 1] - dR[1, 0]])  # JPL left error, small angle
        ep = v[4:7] - p
        P6 = be.get_marginal_covariance([be.handle_imu()])[:6, :6]
        rec["err_ori_deg"] = float(np.degrees(np.linalg.norm(th)))
        rec["err_pos"] = float(np.linalg.norm(ep))
        rec["nees_ori"] = float(th @ np.linalg.solve(P6[:3, :3], th))
        rec["nees_pos"] = float(ep @ np.linalg.solve(P6[3:6, 3:6], ep))
        rec["imu"] = v.copy()
        rec["N"] = be.cov_rows()
        # what sim_save_total_state_to_file writes (ROSVisualizerHelper.cpp:152-300): estimate, 1-sigma of the marginals, simulated truth
        calib, _ = be.var_get(be.handle_calib())
        intr, _ = be.var_get(be.handle_intrinsics())
        Pm = be.get_marginal_covariance([be.handle_imu(), be.handle_intrinsics(), be.handle_calib()])
        rec["calib"], rec["intr"] = np.asarray(calib[:7]).copy(), np.asarray(intr[:8]).copy()
        rec["std"] = np.sqrt(np.maximum(np.diag(Pm), 0.0))  # [imu 15 | intrinsics 8 | extrinsics 6]
        rec["gt"] = np.concatenate([jpl.rot_2_quat(R), p, vel, np.zeros(3), np.zeros(3)])
        self.frames.append(rec)


class _null(object):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def run(backend, n_frames=100, seed=0, max_clones=30, n_feats=60, use_planes=True, gate_ctx=None, keep_cov_every=0, t0=0.5, fit_planes=False):
    """Run the loop for n_frames camera frames; returns (loop, list of (frame index, covariance) snapshots)."""
    sim = RoomSimulator(seed=seed, n_feats=n_feats)
    loop = VioLoop(backend, sim, max_clones=max_clones, gate_ctx=gate_ctx, use_planes=use_planes, fit_planes=fit_planes)
    loop.initialize_with_gt(t0)
    for (ti, wm, am) in sim.imu_until(t0):
        backend.feed_imu(ti, wm, am)
    covs = []
    for k in range(1, n_frames + 1):
        t = t0 + k * sim.cam_dt
        loop._feed_imu(t)
        obs = sim.camera_frame(t)
        loop.step(t, obs)
        if keep_cov_every and k % keep_cov_every == 0:
            covs.append((k, backend.cov()))
    return loop, covs


TIMING_HEADER = "# timestamp (sec),tracking,propagation,plane init,msckf update,re-tri & marg,total"  # VioManager.cpp:110-118 (no SLAM columns: max_slam = 0)


def timing_csv(loop):
    """record_timing_filepath (VioManager.cpp:110-118 header, :911-928 rows: timestamp with 15 decimals, stage times with 5)"""
    lines = [TIMING_HEADER]
    for r in loop.frames:
        tot = r["propagation"] + r["plane_init"] + r["msckf"] + r["marg"]
        lines.append("%.15f,%.5f,%.5f,%.5f,%.5f,%.5f,%.5f" % (r["t"], r.get("front_end", 0.0), r["propagation"], r["plane_init"], r["msckf"], r["marg"],
                                                              tot + r.get("front_end", 0.0)))
    return "\n".join(lines) + "\n"


def _fix(v, prec):
    return ("%." + str(prec) + "f") % v


def state_files(loop):
    """The three text files of sim_save_total_state_to_file (ROSVisualizerHelper.cpp:152-300), one line per frame, fields separated and
    terminated by a blank like the reference's stream writes:
      estimate: t(5 decimals) q_GtoI(4) p v bg ba (6 decimals) t_off(7) num_cameras(0) [intrinsics(8) q_ItoC(4) p_IinC(3)] (6)
      1-sigma : t(5) std of [theta p v bg ba](15) std t_off num_cameras [std intrinsics(8) std extrinsics(6)]   (zeros when not calibrated)
      truth   : same layout as the estimate, from the simulator (biases are zero in this simulator, time offset 0)
    Returns (est, std, gt) strings that ov_eval's error_simulation / timing tools read unchanged."""
    sim = loop.sim
    est, std, gt = [], [], []
    for r in loop.frames:
        t = _fix(r["t"], 5)
        est.append(" ".join([t] + [_fix(x, 6) for x in r["imu"][:16]] + [_fix(0.0, 7), "1"] + [_fix(x, 6) for x in r["intr"]] +
                            [_fix(x, 6) for x in r["calib"]]) + " ")
        sd = r["std"]
        n_intr = 8 if len(sd) >= 15 + 8 else 0
        n_ext = 6 if len(sd) >= 15 + n_intr + 6 else 0
        intr_sd = list(sd[15:15 + n_intr]) if n_intr else [0.0] * 8
        ext_sd = list(sd[15 + n_intr:15 + n_intr + n_ext]) if n_ext else [0.0] * 6
        std.append(" ".join([t] + [_fix(x, 6) for x in sd[:15]] + [_fix(0.0, 6), "1"] + [_fix(x, 6) for x in intr_sd] + [_fix(x, 6) for x in ext_sd]) + " ")
        true_calib = np.concatenate([jpl.rot_2_quat(sim.R_ItoC), sim.p_IinC])
        gt.append(" ".join([t] + [_fix(x, 6) for x in r["gt"]] + [_fix(0.0, 7), "1"] + [_fix(x, 6) for x in sim.cam] + [_fix(x, 6) for x in true_calib]) + " ")
    return "\n".join(est) + "\n", "\n".join(std) + "\n", "\n".join(gt) + "\n"
