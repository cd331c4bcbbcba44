"""TEST INFRASTRUCTURE: ctypes mirror of oracle/liboracle.so with the same method names as ov_plane_b200.api.Context, so a
parity test feeds one scenario to both and compares.  Never imported by the product package."""
import ctypes as C
import os
import subprocess
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ODIR = os.path.join(_ROOT, "oracle")
_LIB = os.path.join(_ODIR, "liboracle.so")

KIND_VEC, KIND_POSE, KIND_IMU, KIND_LANDMARK = 0, 1, 2, 3


def _load():
    srcs = [os.path.join(_ODIR, f) for f in ("oracle.hpp", "oracle_planefit.hpp", "oracle_capi.cpp")]
    if (not os.path.exists(_LIB)) or any(os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-C", _ODIR, "CXX=g++"])
    lib = C.CDLL(_LIB)
    lib.orc_create.restype = C.c_void_p
    lib.orc_last_error.restype = C.c_char_p
    lib.orc_get_timestamp.restype = C.c_double
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


class OracleError(RuntimeError):
    pass


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _cm(a):
    return np.asfortranarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def load_variant(path):
    """A differently compiled build of the same oracle sources (tools/oracle_sensitivity.py: FMA-contracted build)."""
    L = C.CDLL(path)
    L.orc_create.restype = C.c_void_p
    L.orc_last_error.restype = C.c_char_p
    L.orc_get_timestamp.restype = C.c_double
    return L


_PROBE_T = C.CFUNCTYPE(C.c_double, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double))


class GaugeProbe(object):
    """Context manager around the oracle's GaugeProbe instrumentation (oracle.hpp): splits every gated plane chi2 into its
    well-defined part and the squared projection of the compressed residual onto the left null space of the compressed Jacobian
    (rows the reference keeps although their Jacobian part is round-off).  `gate_without=True` makes the oracle gate on the
    well-defined part (what the CUDA path computes); `records` holds (rows, cols, rank, junk, gap) per gated system."""

    def __init__(self, gate_without, variant=None, rank_tol=1e-7):
        self.lib = variant if variant is not None else lib()
        self.gate_without = gate_without
        self.rank_tol = rank_tol
        self.records = []
        self._cb = _PROBE_T(self._probe)

    def _probe(self, rows, cols, Hp, zp):
        H = np.ctypeslib.as_array(Hp, (cols, rows)).T
        z = np.ctypeslib.as_array(zp, (rows,))
        if rows == 0 or cols == 0:
            return 0.0
        U, sv, _ = np.linalg.svd(H, full_matrices=True)
        rank = int((sv > self.rank_tol * sv[0]).sum())
        zn = U[:, rank:].T @ z
        junk = float(zn @ zn)
        gap = (sv[rank - 1] / sv[0], (sv[rank] / sv[0]) if rank < len(sv) else 0.0)
        self.records.append((rows, cols, rank, junk, gap))
        return junk

    def __enter__(self):
        self.lib.orc_set_gauge_probe(self._cb, int(self.gate_without))
        return self

    def __exit__(self, *a):
        self.lib.orc_set_gauge_probe(_PROBE_T(), 0)
        return False

    def junk(self):
        return np.array([r[3] for r in self.records])


class OracleContext(object):
    def __init__(self, o, variant=None, **_unused):
        self.lib = variant if variant is not None else lib()
        self.h = C.c_void_p(self.lib.orc_create(int(o["do_fej"]), int(o["use_rk4_integration"]), int(o["imu_avg"]),
                                                int(o["do_calib_camera_pose"]), int(o["do_calib_camera_intrinsics"]),
                                                int(o["do_calib_camera_timeoffset"]), int(o["max_clone_size"]),
                                                C.c_double(o["sigma_constraint"]), C.c_double(o["const_init_multi"]),
                                                C.c_double(o["const_init_chi2"])))
        self.lib.orc_set_plane_merge_options(self.h, C.c_double(o["sigma_plane_merge"]), C.c_double(o["plane_merge_chi2"]),
                                             C.c_double(o["plane_merge_deg_max"]))

    def close(self):
        if self.h:
            self.lib.orc_destroy(self.h)
            self.h = None

    def _ck(self, st):
        if st != 0:
            raise OracleError("oracle status %d: %s" % (st, self.lib.orc_last_error(self.h).decode()))

    def set_chi2_table(self, q):
        q = _f64(q)
        self.lib.orc_set_chi2_table(self.h, _p(q), len(q))

    def cov_rows(self):
        return self.lib.orc_cov_rows(self.h)

    def cov(self):
        n = self.cov_rows()
        out = np.zeros((n, n), order="F")
        self.lib.orc_get_cov(self.h, _p(out))
        return out

    def cov_upload(self, P):
        P = _cm(P)
        self._ck(self.lib.orc_set_cov(self.h, _p(P), P.shape[0]))

    def handle_imu(self):
        return self.lib.orc_handle_imu(self.h)

    def handle_dt(self):
        return self.lib.orc_handle_dt(self.h)

    def handle_calib(self):
        return self.lib.orc_handle_calib(self.h)

    def handle_intrinsics(self):
        return self.lib.orc_handle_intr(self.h)

    def var_id(self, h):
        return self.lib.orc_var_id(self.h, h)

    def var_size(self, h):
        return self.lib.orc_var_size(self.h, h)

    def var_set(self, h, value, fej=None):
        v = _f64(value)
        f = _f64(fej) if fej is not None else None
        self.lib.orc_var_set(self.h, h, _p(v), _p(f))

    def var_get(self, h):
        n = self.lib.orc_var_nvalue(self.h, h)
        v, f = np.zeros(n), np.zeros(n)
        self.lib.orc_var_get(self.h, h, _p(v), _p(f))
        return v, f

    def variable_order(self):
        n = self.lib.orc_num_variables(self.h)
        o = np.zeros(n, dtype=np.int32)
        self.lib.orc_variable_order(self.h, _p(o))
        return o.tolist()

    def set_timestamp(self, t):
        self.lib.orc_set_timestamp(self.h, C.c_double(t))

    def get_timestamp(self):
        return self.lib.orc_get_timestamp(self.h)

    def add_clone_raw(self, t, v, f):
        v, f = _f64(v), _f64(f)
        return self.lib.orc_add_clone_raw(self.h, C.c_double(t), _p(v), _p(f))

    def add_plane_raw(self, pid, v, f):
        v, f = _f64(v), _f64(f)
        return self.lib.orc_add_plane_raw(self.h, C.c_longlong(int(pid)), _p(v), _p(f))

    def add_slam_raw(self, fid, v, f):
        v, f = _f64(v), _f64(f)
        return self.lib.orc_add_slam_raw(self.h, C.c_longlong(int(fid)), _p(v), _p(f))

    def plane_handle(self, pid):
        return self.lib.orc_plane_handle(self.h, C.c_longlong(int(pid)))

    # ---- StateHelper ----
    def set_initial_covariance(self, cov, handles):
        cov, hs = _cm(cov), _i32(handles)
        self._ck(self.lib.orc_set_initial_covariance(self.h, _p(cov), cov.shape[0], _p(hs), len(hs)))

    def get_marginal_covariance(self, handles):
        hs = _i32(handles)
        n = sum(self.var_size(int(h)) for h in hs)
        out = np.zeros((n, n), order="F")
        self._ck(self.lib.orc_get_marginal_covariance(self.h, _p(hs), len(hs), _p(out)))
        return out

    def ekf_propagation(self, new_handles, old_handles, Phi, Q):
        nh, oh, Phi, Q = _i32(new_handles), _i32(old_handles), _cm(Phi), _cm(Q)
        self._ck(self.lib.orc_ekf_propagation(self.h, _p(nh), len(nh), _p(oh), len(oh), _p(Phi), Phi.shape[0], Phi.shape[1], _p(Q)))

    def ekf_update(self, handles, H, res, Rdiag=None):
        hs, H, res = _i32(handles), _cm(H), _f64(res)
        R = _f64(Rdiag) if Rdiag is not None else None
        self._ck(self.lib.orc_ekf_update(self.h, _p(hs), len(hs), _p(H), H.shape[0], _p(res), _p(R)))

    def marginalize(self, h):
        self._ck(self.lib.orc_marginalize(self.h, h))

    def marginalize_old_clone(self):
        self._ck(self.lib.orc_marginalize_old_clone(self.h))

    def augment_clone(self, t, last_w):
        w, nh = _f64(last_w), C.c_int(-1)
        self._ck(self.lib.orc_augment_clone(self.h, C.c_double(t), _p(w), C.byref(nh)))
        return nh.value

    def initialize(self, kind, value, fej, tag, handles, H_R, H_L, res, sigma2, chi2_mult, do_update=True):
        v, f, hs = _f64(value), _f64(fej), _i32(handles)
        H_R, H_L, res = _cm(H_R), _cm(H_L), _f64(res)
        acc, nh = C.c_int(0), C.c_int(-1)
        self._ck(self.lib.orc_initialize(self.h, kind, len(v), _p(v), _p(f), C.c_longlong(int(tag)), _p(hs), len(hs), _p(H_R), _p(H_L),
                                         _p(res), H_R.shape[0], C.c_double(sigma2), C.c_double(chi2_mult), int(do_update), C.byref(acc),
                                         C.byref(nh)))
        return bool(acc.value), nh.value

    def initialize_invertible(self, kind, value, fej, tag, handles, H_R, H_L, res, sigma2):
        v, f, hs = _f64(value), _f64(fej), _i32(handles)
        H_R, H_L, res = _cm(H_R), _cm(H_L), _f64(res)
        nh = C.c_int(-1)
        self._ck(self.lib.orc_initialize_invertible(self.h, kind, len(v), _p(v), _p(f), C.c_longlong(int(tag)), _p(hs), len(hs), _p(H_R), _p(H_L),
                                                    _p(res), C.c_double(sigma2), C.byref(nh)))
        return nh.value

    def merge_planes_and_marginalize(self, feat2plane, plane2oldplane):
        ff = np.array(list(feat2plane.keys()), dtype=np.int64)
        fp = np.array(list(feat2plane.values()), dtype=np.int64)
        mn, mo = [], []
        for k, olds in plane2oldplane.items():
            for o in olds:
                mn.append(k)
                mo.append(o)
        mn, mo = np.array(mn, dtype=np.int64), np.array(mo, dtype=np.int64)
        self._ck(self.lib.orc_merge_planes_and_marginalize(self.h, _p(ff), _p(fp), len(ff), _p(mn), _p(mo), len(mn)))

    # ---- helpers ----
    def feature_jacobian_full(self, clone_handles, uv, p_FinG, p_FinG_fej, planeid, cp, cp_fej, sigma_px, sigma_c):
        ch, uv = _i32(clone_handles), np.ascontiguousarray(uv, dtype=np.float32)
        m = len(ch)
        rows_cap, cols_cap = 3 * m + 1, 14 + 6 * m + 3
        Hf, Hx, res = np.zeros(rows_cap * 6), np.zeros(rows_cap * cols_cap), np.zeros(rows_cap)
        xo = np.zeros(m + 3, dtype=np.int32)
        hfc, hxc, rows, xon = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        pf, pff = _f64(p_FinG), _f64(p_FinG_fej)
        cpv = _f64(cp) if cp is not None else np.zeros(3)
        cpf = _f64(cp_fej) if cp_fej is not None else np.zeros(3)
        self._ck(self.lib.orc_feature_jacobian_full(self.h, m, _p(ch), _p(uv), _p(pf), _p(pff), C.c_longlong(int(planeid)), _p(cpv), _p(cpf),
                                                    C.c_double(sigma_px), C.c_double(sigma_c), _p(Hf), C.byref(hfc), _p(Hx), C.byref(hxc),
                                                    _p(res), C.byref(rows), _p(xo), C.byref(xon)))
        r = rows.value
        return (Hf[:r * hfc.value].reshape((r, hfc.value), order="F").copy(), Hx[:r * hxc.value].reshape((r, hxc.value), order="F").copy(),
                res[:r].copy(), xo[:xon.value].tolist())

    def feature_jacobian_full_rep(self, clone_handles, uv, representation, anchor_clone_handle, p_F, p_F_fej, sigma_px):
        ch, uv = _i32(clone_handles), np.ascontiguousarray(uv, dtype=np.float32)
        m = len(ch)
        rows_cap, cols_cap = 2 * m, 14 + 6 * (m + 1)
        Hf, Hx, res = np.zeros(rows_cap * 3), np.zeros(rows_cap * cols_cap), np.zeros(rows_cap)
        xo = np.zeros(m + 4, dtype=np.int32)
        hfc, hxc, rows, xon = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        pf, pff = _f64(p_F), _f64(p_F_fej)
        self._ck(self.lib.orc_feature_jacobian_full_rep(self.h, m, _p(ch), _p(uv), int(representation), int(anchor_clone_handle), _p(pf), _p(pff),
                                                        C.c_double(sigma_px), _p(Hf), C.byref(hfc), _p(Hx), C.byref(hxc), _p(res), C.byref(rows),
                                                        _p(xo), C.byref(xon)))
        r = rows.value
        return (Hf[:r * hfc.value].reshape((r, hfc.value), order="F").copy(), Hx[:r * hxc.value].reshape((r, hxc.value), order="F").copy(),
                res[:r].copy(), xo[:xon.value].tolist())

    def slam_set_representation(self, featid, representation, anchor_clone_handle=-1):
        self._ck(self.lib.orc_slam_set_representation(self.h, C.c_longlong(int(featid)), int(representation), int(anchor_clone_handle)))

    def slam_get_representation(self, featid):
        rep, anc = C.c_int(-1), C.c_int(-1)
        self._ck(self.lib.orc_slam_get_representation(self.h, C.c_longlong(int(featid)), C.byref(rep), C.byref(anc)))
        return rep.value, anc.value

    def slam_perform_anchor_change(self, featid, new_anchor_clone_handle):
        self._ck(self.lib.orc_slam_perform_anchor_change(self.h, C.c_longlong(int(featid)), int(new_anchor_clone_handle)))

    def slam_change_anchors(self):
        n = C.c_int(0)
        self._ck(self.lib.orc_slam_change_anchors(self.h, C.byref(n)))
        return n.value

    def nullspace_project_inplace(self, H_f, H_x, res, H_cp=None):
        H_f, H_x, res = _cm(H_f).copy(order="F"), _cm(H_x).copy(order="F"), _f64(res).copy()
        rows, ro = H_f.shape[0], C.c_int()
        if H_cp is None:
            self.lib.orc_nullspace_project_inplace(_p(H_f), H_f.shape[1], _p(H_x), H_x.shape[1], _p(res), rows, C.byref(ro))
            r = ro.value
            return H_x.ravel(order="F")[:r * H_x.shape[1]].reshape((r, H_x.shape[1]), order="F").copy(), res[:r].copy()
        H_cp = _cm(H_cp).copy(order="F")
        self.lib.orc_plane_nullspace_project_inplace(_p(H_f), H_f.shape[1], _p(H_x), H_x.shape[1], _p(H_cp), _p(res), rows, C.byref(ro))
        r = ro.value
        return (H_x.ravel(order="F")[:r * H_x.shape[1]].reshape((r, H_x.shape[1]), order="F").copy(),
                H_cp.ravel(order="F")[:r * 3].reshape((r, 3), order="F").copy(), res[:r].copy())

    def measurement_compress_inplace(self, H_x, res, H_cp=None):
        H_x, res = _cm(H_x).copy(order="F"), _f64(res).copy()
        rows, cols, ro = H_x.shape[0], H_x.shape[1], C.c_int()
        if H_cp is None:
            self.lib.orc_measurement_compress_inplace(_p(H_x), cols, _p(res), rows, C.byref(ro))
            r = ro.value
            return H_x.ravel(order="F")[:r * cols].reshape((r, cols), order="F").copy(), res[:r].copy()
        H_cp = _cm(H_cp).copy(order="F")
        self.lib.orc_plane_measurement_compress_inplace(_p(H_x), cols, _p(H_cp), _p(res), rows, C.byref(ro))
        r = ro.value
        return (H_x.ravel(order="F")[:r * cols].reshape((r, cols), order="F").copy(), H_cp.ravel(order="F")[:r * 3].reshape((r, 3), order="F").copy(),
                res[:r].copy())

    # ---- UpdaterMSCKF ----
    def msckf_update(self, b, sigma_pix=1.0, chi2_mult=1.0, timers=None):
        F, npl = int(b["F"]), len(b["plane_ids"])
        fs, fc = np.zeros(F, dtype=np.int32), np.zeros(F)
        ps, pc = np.zeros(max(1, npl), dtype=np.int32), np.zeros(max(1, npl))
        hx, hxn = np.zeros(4096, dtype=np.int32), C.c_int(0)
        t4 = np.zeros(4)
        arrs = [np.ascontiguousarray(b[k]) for k in ("meas_offset", "meas_clone", "uv", "p_FinG", "p_FinG_original", "featid", "planeid",
                                                      "plane_ids", "plane_cp")]
        self._ck(self.lib.orc_msckf_update(self.h, F, _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), _p(arrs[3]), _p(arrs[4]), _p(arrs[5]), _p(arrs[6]),
                                           npl, _p(arrs[7]), _p(arrs[8]), C.c_double(sigma_pix), C.c_double(chi2_mult), _p(fs), _p(fc), _p(ps),
                                           _p(pc), _p(hx), C.byref(hxn), _p(t4)))
        if timers is not None:
            timers[:] = t4
        return dict(feat_status=fs, feat_chi2=fc, plane_status=ps[:npl], plane_chi2=pc[:npl], hx_order=hx[:hxn.value].tolist())

    def plane_init(self, b, sigma_pix=1.0):
        F, npl = int(b["F"]), len(b["plane_ids"])
        ps, nh = np.zeros(max(1, npl), dtype=np.int32), np.zeros(max(1, npl), dtype=np.int32)
        arrs = [np.ascontiguousarray(b[k]) for k in ("meas_offset", "meas_clone", "uv", "p_FinG", "featid", "planeid", "plane_ids", "plane_cp")]
        self._ck(self.lib.orc_plane_init(self.h, F, _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), _p(arrs[3]), _p(arrs[4]), _p(arrs[5]), npl,
                                         _p(arrs[6]), _p(arrs[7]), C.c_double(sigma_pix), _p(ps), _p(nh)))
        return dict(plane_status=ps[:npl], new_handles=nh[:npl])

    def slam_update(self, b, sigma_pix=1.0, chi2_mult=1.0):
        F = int(b["F"])
        fs, fc = np.zeros(F, dtype=np.int32), np.zeros(F)
        arrs = [np.ascontiguousarray(b[k]) for k in ("meas_offset", "meas_clone", "uv", "featid", "planeid")]
        self._ck(self.lib.orc_slam_update(self.h, F, _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), _p(arrs[3]), _p(arrs[4]), C.c_double(sigma_pix),
                                          C.c_double(chi2_mult), _p(fs), _p(fc)))
        return dict(feat_status=fs, feat_chi2=fc)

    def slam_delayed_init(self, b, sigma_pix=1.0, chi2_mult=1.0):
        F = int(b["F"])
        fs, nh = np.zeros(F, dtype=np.int32), np.zeros(F, dtype=np.int32)
        arrs = [np.ascontiguousarray(b[k]) for k in ("meas_offset", "meas_clone", "uv", "p_FinG", "p_FinG_original", "featid", "planeid")]
        self._ck(self.lib.orc_slam_delayed_init(self.h, F, _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), _p(arrs[3]), _p(arrs[4]), _p(arrs[5]),
                                                _p(arrs[6]), C.c_double(sigma_pix), C.c_double(chi2_mult), _p(fs), _p(nh)))
        return dict(feat_status=fs, new_handles=nh)

    def marginalize_slam(self):
        self._ck(self.lib.orc_marginalize_slam(self.h))

    def slam_handle(self, featid):
        return self.lib.orc_slam_handle(self.h, C.c_longlong(int(featid)))

    def slam_should_marg(self, featid):
        return self.lib.orc_slam_should_marg(self.h, C.c_longlong(int(featid)))

    # ---- Propagator ----
    def propagator_set_noise(self, sigma_w, sigma_wb, sigma_a, sigma_ab, gravity_mag=9.81):
        self.lib.orc_prop_set(self.h, C.c_double(sigma_w), C.c_double(sigma_wb), C.c_double(sigma_a), C.c_double(sigma_ab),
                              C.c_double(gravity_mag))

    def feed_imu(self, t, wm, am):
        w, a = _f64(wm), _f64(am)
        self.lib.orc_prop_feed_imu(self.h, C.c_double(t), _p(w), _p(a))

    def triangulate_features(self, meas_offset, meas_clone, uv_norm):
        mo, mc = _i32(meas_offset), _i32(meas_clone)
        uvn = np.ascontiguousarray(uv_norm, dtype=np.float32)
        F = len(mo) - 1
        pf, st = np.zeros((max(1, F), 3)), np.zeros(max(1, F), dtype=np.int32)
        self._ck(self.lib.orc_triangulate_features(self.h, F, _p(mo), _p(mc), _p(uvn), _p(pf), _p(st)))
        return pf[:F], st[:F]

    # ---- PlaneFitting ----
    def plane_fitting(self, feat_offset, p_FinG, min_inlier_num, max_cond, shuffle_kind=0):
        fo, pf = _i32(feat_offset), _f64(p_FinG).reshape(-1, 3)
        nP, Ft = len(fo) - 1, int(fo[-1])
        st, ab, inl = np.zeros(nP, dtype=np.int32), np.zeros((nP, 4)), np.zeros(max(1, Ft), dtype=np.int32)
        for p in range(nP):
            a, b = int(fo[p]), int(fo[p + 1])
            pts, il, ok, abcd = np.ascontiguousarray(pf[a:b]), np.zeros(max(1, b - a), dtype=np.int32), C.c_int(0), np.zeros(4)
            self.lib.orc_plane_fitting(b - a, _p(pts), int(min_inlier_num), C.c_double(max_cond), int(shuffle_kind), _p(abcd), _p(il), C.byref(ok))
            st[p] = ok.value
            ab[p] = abcd if ok.value else 0.0
            inl[a:b] = il[:b - a] if ok.value else 0
        return st, ab, inl[:Ft]

    def optimize_plane(self, feat_offset, meas_offset, meas_clone, uv_norm, p_FinG, cp_inG, fix_plane, sigma_px_norm, sigma_c, max_num_iterations=0):
        fo, mo, mc = _i32(feat_offset), _i32(meas_offset), _i32(meas_clone)
        uvn = np.ascontiguousarray(uv_norm, dtype=np.float32).reshape(-1, 2)
        pf, cp, fx = _f64(p_FinG).reshape(-1, 3), _f64(cp_inG).reshape(-1, 3), _i32(fix_plane)
        nP, Ft = len(fo) - 1, int(fo[-1])
        po, co = pf.copy(), cp.copy()
        inl, st, info = np.zeros(max(1, Ft), dtype=np.int32), np.zeros(nP, dtype=np.int32), np.zeros((nP, 5))
        for p in range(nP):
            a, b = int(fo[p]), int(fo[p + 1])
            F = b - a
            m0, m1 = int(mo[a]), int(mo[b])
            lmo = np.ascontiguousarray(mo[a:b + 1] - m0, dtype=np.int32)
            lmc, luv = np.ascontiguousarray(mc[m0:m1]), np.ascontiguousarray(uvn[m0:m1])
            lp, lcp = np.ascontiguousarray(pf[a:b]), np.ascontiguousarray(cp[p])
            op, oc, il, ok, inf = np.zeros((max(1, F), 3)), np.zeros(3), np.zeros(max(1, F), dtype=np.int32), C.c_int(0), np.zeros(5)
            self._ck(self.lib.orc_optimize_plane(self.h, F, _p(lmo), _p(lmc), _p(luv), _p(lp), _p(lcp), C.c_double(sigma_px_norm),
                                                 C.c_double(sigma_c), int(fx[p]), int(max_num_iterations), _p(op), _p(oc), _p(il), C.byref(ok), _p(inf)))
            st[p], info[p] = ok.value, inf
            po[a:b], co[p], inl[a:b] = op[:F], oc, il[:F]
        return st, po, co, inl[:Ft], info

    def optimize_plane_cost(self, meas_offset, meas_clone, uv_norm, p_FinG, cp_inG, fix_plane, sigma_px_norm, sigma_c):
        """robustified cost 1/2 sum rho(s) of ONE plane's refinement problem at the given point"""
        mo, mc = _i32(meas_offset), _i32(meas_clone)
        uvn = np.ascontiguousarray(uv_norm, dtype=np.float32)
        pf, cp = _f64(p_FinG).reshape(-1, 3), _f64(cp_inG)
        cost = C.c_double(0.0)
        self._ck(self.lib.orc_optimize_plane_cost(self.h, len(mo) - 1, _p(mo), _p(mc), _p(uvn), _p(pf), _p(cp), C.c_double(sigma_px_norm),
                                                  C.c_double(sigma_c), int(fix_plane), C.byref(cost)))
        return cost.value

    # ---- UpdaterZeroVelocity ----
    def zupt_feed_imu(self, t, wm, am):
        w, a = _f64(wm), _f64(am)
        self.lib.orc_zupt_feed_imu(self.h, C.c_double(t), _p(w), _p(a))

    def zupt_try_update(self, t, average_disparity, num_features, gravity_mag=9.81, max_velocity=1.0, noise_multiplier=1.0, max_disparity=1.0,
                        chi2_mult=1.0, noises=(1.6968e-04, 1.9393e-05, 2.0e-3, 3.0e-3)):
        self.lib.orc_zupt_set(self.h, *[C.c_double(x) for x in noises], C.c_double(gravity_mag), C.c_double(max_velocity),
                              C.c_double(noise_multiplier), C.c_double(max_disparity), C.c_double(chi2_mult))
        acc, chi = C.c_int(0), C.c_double(0.0)
        self._ck(self.lib.orc_zupt_try_update(self.h, C.c_double(t), C.c_double(average_disparity), int(num_features), C.byref(acc), C.byref(chi)))
        return bool(acc.value), chi.value

    def fast_state_propagate(self, t):
        sp, cv, ok = np.zeros(13), np.zeros((12, 12), order="F"), C.c_int(0)
        self._ck(self.lib.orc_prop_fast_state_propagate(self.h, C.c_double(t), _p(sp), _p(cv), C.byref(ok)))
        return (sp, cv) if ok.value else None

    def propagate_and_clone(self, t):
        Phi, Q, nh = np.zeros((15, 15), order="F"), np.zeros((15, 15), order="F"), C.c_int(-1)
        self._ck(self.lib.orc_prop_propagate_and_clone(self.h, C.c_double(t), _p(Phi), _p(Q), C.byref(nh)))
        return nh.value, Phi, Q
