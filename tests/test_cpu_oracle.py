"""Pins the CPU oracle (oracle/oracle.hpp).  The reference ships no tests or golden vectors, so the
oracle is pinned by (i) mathematical properties that define correctness independently of any implementation (numerical
derivatives, orthogonality, algebraic identities of the Kalman update), (ii) a NumPy/LAPACK second opinion, (iii) committed
golden vectors (tests/golden/, regression pin of the oracle itself; regenerate with tests/golden/make_golden.py)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import oracle_backend as ob
from ov_plane_b200 import jpl, synth

L = ob.lib()


def _v(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def q2R(q):
    R = np.zeros((3, 3), order="F")
    L.orc_quat_2_Rot(_v(q).ctypes.data_as(C.c_void_p), R.ctypes.data_as(C.c_void_p))
    return np.array(R)


def test_quaternion_ops():
    rng = np.random.RandomState(0)
    for _ in range(20):
        q = rng.randn(4)
        q /= np.linalg.norm(q)
        R = q2R(q)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-14) and abs(np.linalg.det(R) - 1) < 1e-13
        assert np.allclose(R, jpl.quat_2_Rot(q), atol=1e-15)
        q2 = np.zeros(4)
        L.orc_rot_2_quat(np.asfortranarray(R).ctypes.data_as(C.c_void_p), q2.ctypes.data_as(C.c_void_p))
        assert np.allclose(q2, q if q[3] >= 0 else -q, atol=1e-12)
        p = rng.randn(4)
        p /= np.linalg.norm(p)
        qp = np.zeros(4)
        L.orc_quat_multiply(_v(q).ctypes.data_as(C.c_void_p), _v(p).ctypes.data_as(C.c_void_p), qp.ctypes.data_as(C.c_void_p))
        assert np.allclose(q2R(qp), q2R(q) @ q2R(p), atol=1e-13)  # JPL: R(q (x) p) = R(q) R(p)
        assert qp[3] >= 0 and abs(np.linalg.norm(qp) - 1) < 1e-14


def test_exp_so3_and_Jr():
    from scipy.linalg import expm
    rng = np.random.RandomState(1)
    for sc in (1e-9, 1e-3, 0.5, 2.0):
        w = sc * rng.randn(3)
        R = np.zeros((3, 3), order="F")
        L.orc_exp_so3(_v(w).ctypes.data_as(C.c_void_p), R.ctypes.data_as(C.c_void_p))
        assert np.allclose(R, expm(jpl.skew(w)), atol=1e-12)
        J = np.zeros((3, 3), order="F")
        L.orc_Jr_so3(_v(w).ctypes.data_as(C.c_void_p), J.ctypes.data_as(C.c_void_p))
        d = 1e-6 * rng.randn(3)
        lhs = expm(jpl.skew(w + d))
        rhs = expm(jpl.skew(w)) @ expm(jpl.skew(J @ d))  # right Jacobian: exp(w + d) ~ exp(w) exp(Jr d)
        assert np.allclose(lhs, rhs, atol=1e-10)


def test_jpl_update_is_left_multiplicative():
    """R(q (+) dth) = exp(-[dth x]) R(q): the convention the clone Jacobian R_ItoC [p_FinIi x] assumes (UpdaterHelper.cpp:400)."""
    from scipy.linalg import expm
    rng = np.random.RandomState(2)
    for kind, nv in ((ob.KIND_POSE, 7), (ob.KIND_IMU, 16)):
        val = rng.randn(nv)
        val[:4] /= np.linalg.norm(val[:4])
        dx = 1e-2 * rng.randn(15)
        new = val.copy()
        L.orc_var_update(kind, new.ctypes.data_as(C.c_void_p), dx.ctypes.data_as(C.c_void_p))
        assert np.allclose(q2R(new[:4]), expm(-jpl.skew(dx[:3])) @ q2R(val[:4]), atol=1e-6 * 1e-2)
        assert np.allclose(new[4:7], val[4:7] + dx[3:6])
        if kind == ob.KIND_IMU:
            assert np.allclose(new[7:16], val[7:16] + dx[6:15])


def test_radtan_jacobian_numeric():
    cam = synth.EUROC_CAM.copy()
    rng = np.random.RandomState(3)
    for _ in range(10):
        x, y = rng.uniform(-0.6, 0.6, 2)
        uv = np.zeros(2)
        dzn, dze = np.zeros((2, 2), order="F"), np.zeros((2, 8), order="F")
        L.orc_radtan_jacobian(cam.ctypes.data_as(C.c_void_p), C.c_double(x), C.c_double(y), dzn.ctypes.data_as(C.c_void_p),
                              dze.ctypes.data_as(C.c_void_p))

        def f(c, xx, yy):
            o = np.zeros(2)
            L.orc_radtan_distort(_v(c).ctypes.data_as(C.c_void_p), C.c_double(xx), C.c_double(yy), o.ctypes.data_as(C.c_void_p))
            return o
        assert np.allclose(f(cam, x, y), jpl.radtan_distort(cam, x, y))
        h = 1e-6
        num = np.stack([(f(cam, x + h, y) - f(cam, x - h, y)) / (2 * h), (f(cam, x, y + h) - f(cam, x, y - h)) / (2 * h)], axis=1)
        assert np.allclose(dzn, num, rtol=1e-6, atol=1e-6)
        for k in range(8):
            cp, cm = cam.copy(), cam.copy()
            hk = 1e-6 * max(1.0, abs(cam[k]))
            cp[k] += hk
            cm[k] -= hk
            assert np.allclose(dze[:, k], (f(cp, x, y) - f(cm, x, y)) / (2 * hk), rtol=1e-5, atol=1e-6)


def test_make_givens():
    rng = np.random.RandomState(4)
    cases = [(1.0, 0.0), (-2.0, 0.0), (0.0, 3.0), (0.0, -3.0), (0.0, 0.0)] + [tuple(rng.randn(2)) for _ in range(20)]
    for p, q in cases:
        cs = np.zeros(2)
        L.orc_make_givens(C.c_double(p), C.c_double(q), cs.ctypes.data_as(C.c_void_p))
        c, s = cs
        assert abs(c * c + s * s - 1) < 1e-14
        # applyOnTheLeft(0, 1, G.adjoint()): x' = c x - s y ; y' = s x + c y must annihilate the second entry
        assert abs(s * p + c * q) < 1e-14 * max(1.0, abs(p) + abs(q))
        assert abs(abs(c * p - s * q) - np.hypot(p, q)) < 1e-13 * max(1.0, np.hypot(p, q))


def _feature_residual(S, o, ch, f, pf):
    a, b = S.meas_offset[f], S.meas_offset[f + 1]
    idx = S.meas_clone_idx[a:b]
    pid = int(S.planeid[f])
    cp = cpf = None
    if pid:
        cp, cpf = o.var_get(o.plane_handle(pid))
    return o.feature_jacobian_full([ch[i] for i in idx], S.uv[a:b], pf, pf, pid, cp, cpf, 1.0, 0.01)


def test_feature_jacobian_matches_numerical_derivative():
    """Without FEJ the Jacobian must be the derivative of the residual w.r.t. the error state under ov_type's update rules:
    r(x (+) dx, p_f + dp) ~ r(x, p_f) - H_x dx - H_f dp.  Covers bearing rows, extrinsics, intrinsics, clone pose, in-state plane."""
    S = synth.make_scenario("tiny_planes", seed=5)
    opt = dict(S.options)
    opt["do_fej"] = 0
    rng = np.random.RandomState(6)
    checked = 0
    for f in range(S.F):
        o = ob.OracleContext(opt)
        ch = synth.load_scenario_into(o, S)
        pf = S.p_FinG[f].copy()
        Hf, Hx, r0, order = _feature_residual(S, o, ch, f, pf)
        sizes = [o.var_size(h) for h in order]
        for trial in range(2):
            dxs = [1e-6 * rng.randn(s) for s in sizes]
            dp = 1e-6 * rng.randn(3)
            o2 = ob.OracleContext(opt)
            ch2 = synth.load_scenario_into(o2, S)
            # handles are identical across contexts built the same way
            for h, d in zip(order, dxs):
                v, fe = o2.var_get(h)
                kind = ob.KIND_POSE if len(v) == 7 else ob.KIND_VEC
                vv = np.zeros(16)
                vv[:len(v)] = v
                dd = np.zeros(15)
                dd[:len(d)] = d
                if kind == ob.KIND_POSE:
                    L.orc_var_update(kind, vv.ctypes.data_as(C.c_void_p), dd.ctypes.data_as(C.c_void_p))
                else:
                    vv[:len(v)] = v + d
                o2.var_set(h, vv[:len(v)], fe)
            _, _, r1, _ = _feature_residual(S, o2, ch2, f, pf + dp)
            pred = r0 - Hx @ np.concatenate(dxs) - Hf[:, :3] @ dp
            assert np.allclose(r1, pred, atol=5e-9 * max(1.0, np.abs(Hx).max())), np.abs(r1 - pred).max()
            checked += 1
        if checked >= 12:
            break


def _numpy_update(P, cols, H, r, Rd=None):
    Ps = P[np.ix_(cols, cols)]
    S_ = H @ Ps @ H.T + (np.eye(len(r)) if Rd is None else np.diag(Rd))
    K = P[:, cols] @ H.T @ np.linalg.inv(S_)
    return P - K @ H @ P[cols, :], K @ r


def test_ekf_update_against_numpy_and_identities():
    S = synth.make_scenario("tiny_points", seed=7)
    o = ob.OracleContext(S.options)
    ch = synth.load_scenario_into(o, S)
    rng = np.random.RandomState(8)
    hs = [o.handle_calib(), ch[1], ch[4]]
    cols = np.concatenate([np.arange(o.var_id(h), o.var_id(h) + o.var_size(h)) for h in hs])
    H = rng.randn(40, len(cols)) * 30
    r = rng.randn(40)
    P0 = o.cov()
    Pn, dxn = _numpy_update(P0, cols, H, r)
    v0 = {h: o.var_get(h)[0] for h in ch}
    o.ekf_update(hs, H, r)
    P1 = o.cov()
    assert np.linalg.norm(P1 - Pn) / np.linalg.norm(Pn) < 1e-10
    assert np.abs(P1 - P1.T).max() == 0.0
    assert np.linalg.eigvalsh(P1).min() > -1e-12
    # the mean moved by dx under the update rules (position part is additive)
    for h in ch:
        assert np.allclose(o.var_get(h)[0][4:7], v0[h][4:7] + dxn[o.var_id(h) + 3:o.var_id(h) + 6], atol=1e-12)
    # compress-then-update == update with the uncompressed system (measurement_compress_inplace is lossless for (x+, P+))
    o2 = ob.OracleContext(S.options)
    synth.load_scenario_into(o2, S)
    Hc, rc = o2.measurement_compress_inplace(H, r)
    assert Hc.shape == (len(cols), len(cols)) and np.abs(np.tril(Hc, -1)).max() < 1e-9 * np.abs(Hc).max()
    o2.ekf_update(hs, Hc, rc)
    assert np.linalg.norm(o2.cov() - P1) / np.linalg.norm(P1) < 1e-10
    for h in ch:
        assert np.allclose(o2.var_get(h)[0], o.var_get(h)[0], atol=1e-11)
    # column permutation of (H, H_order) leaves the posterior unchanged
    o3 = ob.OracleContext(S.options)
    synth.load_scenario_into(o3, S)
    perm_h = [hs[2], hs[0], hs[1]]
    Hp = np.hstack([H[:, 12:18], H[:, 0:6], H[:, 6:12]])
    o3.ekf_update(perm_h, Hp, r)
    assert np.linalg.norm(o3.cov() - P1) / np.linalg.norm(P1) < 1e-11


def test_nullspace_projection_properties():
    o = ob.OracleContext(synth.make_scenario("tiny_points").options)
    rng = np.random.RandomState(9)
    Hf, Hx, r = rng.randn(20, 3), rng.randn(20, 30), rng.randn(20)
    Ho, ro = o.nullspace_project_inplace(Hf, Hx, r)
    assert Ho.shape == (17, 30)
    Pi = np.eye(20) - Hf @ np.linalg.solve(Hf.T @ Hf, Hf.T)  # projector onto the left nullspace of H_f
    assert np.allclose(Ho.T @ Ho, Hx.T @ Pi @ Hx, atol=1e-10)
    assert np.allclose(Ho.T @ ro, Hx.T @ Pi @ r, atol=1e-10)
    assert abs(ro @ ro - r @ Pi @ r) < 1e-10


def test_propagation_phi_matches_numerical_derivative_of_the_mean():
    """Without FEJ, F = d(x_{k+1} (-) x_hat_{k+1}) / d(error state) for one IMU interval (Propagator.cpp:411-432): checks theta, p, v
    blocks against finite differences of the discrete mean propagation."""
    S = synth.make_scenario("tiny_points", seed=1)
    opt = dict(S.options)
    opt["do_fej"] = 0
    opt["use_rk4_integration"] = 0
    t0 = S.timestamp

    def run(dx):
        o = ob.OracleContext(opt)
        synth.load_scenario_into(o, S)
        v, f = o.var_get(o.handle_imu())
        vv = v.copy()
        L.orc_var_update(ob.KIND_IMU, vv.ctypes.data_as(C.c_void_p), _v(dx).ctypes.data_as(C.c_void_p))
        o.var_set(o.handle_imu(), vv, vv)
        o.var_set(o.handle_dt(), np.zeros(1), np.zeros(1))
        o.propagator_set_noise(1e-4, 1e-5, 1e-3, 1e-3, 9.81)
        o.feed_imu(t0, [0.1, -0.2, 0.15], [0.3, 9.6, 0.5])
        o.feed_imu(t0 + 0.01, [0.1, -0.2, 0.15], [0.3, 9.6, 0.5])
        o.feed_imu(t0 + 0.02, [0.1, -0.2, 0.15], [0.3, 9.6, 0.5])
        _, Phi, Q = o.propagate_and_clone(t0 + 0.01)
        return o.var_get(o.handle_imu())[0], Phi
    x0, Phi = run(np.zeros(15))
    assert np.allclose(Phi[9:15, 9:15], np.eye(6)) and np.allclose(Phi[3:6, 3:6], np.eye(3))
    h = 1e-6
    for k in [0, 1, 2, 3, 6, 7, 9, 12]:
        d = np.zeros(15)
        d[k] = h
        x1, _ = run(d)
        err = np.zeros(15)
        dR = q2R(x1[:4]) @ q2R(x0[:4]).T  # = exp(-[dth x])
        err[0:3] = -np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]) / 2
        err[3:15] = x1[4:16] - x0[4:16]
        assert np.allclose(err / h, Phi[:, k], atol=2e-5), (k, np.abs(err / h - Phi[:, k]).max())


GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name,seed", [("tiny_points", 0), ("tiny_planes", 0), ("small_planes", 1), ("cfg1_euroc_n96", 0), ("cfg2_n256_f200", 0)])
def test_oracle_reproduces_golden_vectors(name, seed, chi2_table):
    g = np.load(os.path.join(GOLD, "%s_s%d.npz" % (name, seed)))
    S = synth.make_scenario(name, seed=seed)
    # (the generator builds P0 with NumPy / BLAS products: bit-identical on the authoring machine, equal to round-off on another CPU)
    assert np.allclose(S.P0, g["P0"], rtol=1e-12, atol=1e-18) and np.array_equal(S.uv, g["uv"]), "scenario generator changed: regenerate the goldens"
    o = ob.OracleContext(S.options)
    o.set_chi2_table(chi2_table)
    ch = synth.load_scenario_into(o, S)
    r = o.msckf_update(synth.feature_batch(S, ch), 1.0, 1.0)
    assert np.array_equal(r["feat_status"], g["feat_status"]) and np.array_equal(r["plane_status"], g["plane_status"])
    assert np.allclose(o.cov(), g["P1"], rtol=0, atol=1e-12 * np.abs(g["P1"]).max())
    assert np.allclose(o.var_get(o.handle_imu())[0], g["imu1"], atol=1e-12)


def test_slam_update_against_numpy_restack(chi2_table):
    """UpdaterSLAM::update of the restatement == an independent numpy re-derivation: per-feature gate on the marginal covariance
    (UpdaterSLAM.cpp:521-532), plane -> no-plane retry (:547-609), one stacked EKF update with R = I (:672-673)."""
    S = synth.make_scenario("small_planes", seed=0)
    orc = ob.OracleContext(S.options)
    orc.set_chi2_table(chi2_table)
    cho = synth.load_scenario_into(orc, S)
    sel = np.arange(24)
    b = synth.feature_batch(S, cho, sel)
    o = orc.slam_delayed_init(b, 1.0, 1.0)
    keep = np.nonzero(o["feat_status"] > 0)[0]
    assert len(keep) >= 16 and orc.cov_rows() == S.N + 3 * len(keep)
    u = synth.feature_batch(S, cho, keep)
    rng = np.random.default_rng(3)
    u["uv"] = u["uv"] + rng.normal(0.0, 0.5, u["uv"].shape).astype(np.float32)
    u["uv"][u["meas_offset"][2]:u["meas_offset"][3]] += 30.0  # one gross outlier
    P0 = orc.cov()
    N = P0.shape[0]
    sc = S.options["sigma_constraint"] if isinstance(S.options, dict) else S.options.sigma_constraint
    # numpy side, from the state BEFORE the update
    Hrows, rrows, expect = [], [], []
    for f in range(len(keep)):
        a, e = u["meas_offset"][f], u["meas_offset"][f + 1]
        lm = orc.slam_handle(u["featid"][f])
        p, pfej = orc.var_get(lm)
        pid = int(u["planeid"][f])
        ph = orc.plane_handle(pid) if pid else -1
        st = 0
        for attempt_plane in ([True, False] if ph >= 0 else [False]):
            cp, cpf = orc.var_get(ph) if attempt_plane else (None, None)
            Hf, Hx, res, xo = orc.feature_jacobian_full(u["meas_clone"][a:e], u["uv"][a:e], p, pfej, pid if attempt_plane else 0, cp, cpf, 1.0, sc)
            Hbig = np.zeros((len(res), N))
            c0 = 0
            for h in xo:
                i, s = orc.var_id(h), orc.var_size(h)
                Hbig[:, i:i + s] += Hx[:, c0:c0 + s]
                c0 += s
            i = orc.var_id(lm)
            Hbig[:, i:i + 3] += Hf[:, :3]
            chi = res @ np.linalg.solve(Hbig @ P0 @ Hbig.T + np.eye(len(res)), res)
            if chi <= chi2_table[len(res)]:
                st = 1 if (attempt_plane or ph < 0) else 3
                Hrows.append(Hbig)
                rrows.append(res)
                break
        expect.append(st)
    got = orc.slam_update(u, 1.0, 1.0)
    assert list(got["feat_status"]) == expect and 0 in expect
    H, r = np.vstack(Hrows), np.concatenate(rrows)
    K = P0 @ H.T @ np.linalg.inv(H @ P0 @ H.T + np.eye(len(r)))
    P1 = P0 - K @ H @ P0
    assert np.abs(orc.cov() - P1).max() < 1e-9 * np.abs(P0).max()
    assert orc.slam_should_marg(u["featid"][2]) == 1


@pytest.mark.parametrize("name,seed,nslam", [("tiny_planes", 0, 6), ("tiny_points", 0, 8)])
def test_oracle_reproduces_slam_golden_vectors(name, seed, nslam, chi2_table):
    sys.path.insert(0, GOLD)
    import make_golden
    g = np.load(os.path.join(GOLD, "slam_%s_s%d.npz" % (name, seed)))
    r = make_golden.slam_case(name, seed, nslam, lambda S: ob.OracleContext(S.options), chi2_table)
    assert np.array_equal(r["init_status"], g["init_status"]) and np.array_equal(r["upd_status"], g["upd_status"])
    assert np.allclose(r["P_init"], g["P_init"], rtol=1e-12, atol=1e-18) and np.allclose(r["P_upd"], g["P_upd"], rtol=1e-10, atol=1e-16)


def test_fast_state_propagate_against_numpy():
    """Propagator::fast_state_propagate of the restatement == a NumPy re-derivation of Propagator.cpp:128-224 (second, independent
    transcription: Jr_so3 by its series definition, Phi / Qd assembled with np.block), and the state is left untouched."""
    S = synth.make_scenario("tiny_points", seed=4)
    orc = ob.OracleContext(S.options)
    synth.load_scenario_into(orc, S)
    sw, swb, sa, sab, grav = 1.6968e-04, 1.9393e-05, 2.0e-3, 3.0e-3, 9.81
    orc.propagator_set_noise(sw, swb, sa, sab, grav)
    rng = np.random.RandomState(5)
    t0 = S.timestamp
    ts, wms, ams = [], [], []
    for k in range(70):
        t = t0 - 0.0323 + 0.0025 * k  # the buffer must cover t0 + t_off (t_off = -13 ms in this scenario)
        wm = np.array([0.3, -0.15, 0.2]) + 0.01 * rng.randn(3)
        am = np.array([0.2, 9.7, 0.4]) + 0.05 * rng.randn(3)
        orc.feed_imu(t, wm, am)
        ts.append(t), wms.append(wm), ams.append(am)
    P_before = orc.cov().copy()
    t1 = t0 + 0.1
    sp, cv = orc.fast_state_propagate(t1)
    assert np.array_equal(orc.cov(), P_before)

    # ---- NumPy side ----
    def Jr(phi):  # right Jacobian of SO(3)
        th = np.linalg.norm(phi)
        K = jpl.skew(phi)
        if th < 1e-7:
            return np.eye(3) - 0.5 * K
        return np.eye(3) - (1 - np.cos(th)) / th**2 * K + (th - np.sin(th)) / th**3 * K @ K

    x = orc.var_get(orc.handle_imu())[0].copy()
    cov = orc.get_marginal_covariance([orc.handle_imu()])
    toff = orc.var_get(orc.handle_dt())[0][0]
    ts, wms, ams = np.array(ts), np.array(wms), np.array(ams)
    a, b = t0 + toff, t1 + toff

    def interp(t):
        i = np.searchsorted(ts, t) - 1
        lam = (t - ts[i]) / (ts[i + 1] - ts[i])
        return (1 - lam) * wms[i] + lam * wms[i + 1], (1 - lam) * ams[i] + lam * ams[i + 1]

    inner = [(t, w, am) for t, w, am in zip(ts, wms, ams) if a < t < b]
    seq = [(a,) + interp(a)] + inner + [(b,) + interp(b)]
    bg, ba = x[10:13], x[13:16]
    g = np.array([0, 0, grav])
    for (ta, wa, aa), (tb, wb, ab) in zip(seq[:-1], seq[1:]):
        dt = tb - ta
        w_hat, a_hat = 0.5 * (wa + wb) - bg, 0.5 * (aa + ab) - ba
        R = jpl.quat_2_Rot(x[0:4])
        E = jpl.exp_so3(-w_hat * dt)
        EJ = -E @ Jr(-w_hat * dt) * dt
        Z, I = np.zeros((3, 3)), np.eye(3)
        F = np.block([[E, Z, Z, EJ, Z],
                      [-0.5 * R.T @ jpl.skew(a_hat * dt * dt), I, I * dt, Z, -0.5 * R.T * dt * dt],
                      [-R.T @ jpl.skew(a_hat * dt), Z, I, Z, -R.T * dt],
                      [Z, Z, Z, I, Z],
                      [Z, Z, Z, Z, I]])
        G = np.block([[EJ, Z, Z, Z], [Z, -0.5 * R.T * dt * dt, Z, Z], [Z, -R.T * dt, Z, Z], [Z, Z, I, Z], [Z, Z, Z, I]])
        Qc = np.diag([sw**2 / dt] * 3 + [sa**2 / dt] * 3 + [swb**2 * dt] * 3 + [sab**2 * dt] * 3)
        Qd = G @ Qc @ G.T
        cov = F @ cov @ F.T + 0.5 * (Qd + Qd.T)
        p, v = x[4:7].copy(), x[7:10].copy()
        x[0:4] = jpl.rot_2_quat(E @ R)
        x[4:7] = p + v * dt + 0.5 * R.T @ a_hat * dt * dt - 0.5 * g * dt * dt
        x[7:10] = v + R.T @ a_hat * dt - g * dt
    Rq = jpl.quat_2_Rot(x[0:4])
    sp_np = np.concatenate([x[0:4], x[4:7], Rq @ x[7:10], 0.5 * (seq[-1][1] + seq[-2][1]) - bg])
    Phi = np.eye(15)
    Phi[6:9, 6:9] = Rq
    c2 = Phi @ cov @ Phi.T
    cv_np = np.zeros((12, 12))
    cv_np[:9, :9] = c2[:9, :9]
    cv_np[9:, 9:] = np.eye(3) * sw**2 / (seq[-1][0] - seq[-2][0])
    assert np.allclose(sp, sp_np, rtol=1e-10, atol=1e-12)
    assert np.abs(cv - cv_np).max() < 1e-10 * np.abs(cv_np).max()
