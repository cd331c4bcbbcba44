"""CPU checks of the oracle's anchored-representation code (it is the checker of tests/test_gpu_anchors.py, and the reference ships no
vectors): get_feature_jacobian_full with a representation against numerical derivatives of its own residual under the ov_type update
rules, and perform_anchor_change through the invariance it must have - the landmark's global position and that position's covariance
are the same before and after the change (exactly, to first order, when first-estimate Jacobians are off)."""
import numpy as np
import pytest

import oracle_backend
from ov_plane_b200 import jpl, synth

REPS = {0: "GLOBAL_3D", 1: "GLOBAL_FULL_INVERSE_DEPTH", 2: "ANCHORED_3D", 3: "ANCHORED_FULL_INVERSE_DEPTH", 4: "ANCHORED_MSCKF_INVERSE_DEPTH"}


def to_lambda(rep, p):
    if rep in (0, 2):
        return p.copy()
    if rep in (1, 3):
        rho = 1 / np.linalg.norm(p)
        return np.array([np.arctan2(p[1], p[0]), np.arccos(rho * p[2]), rho])
    return np.array([p[0] / p[2], p[1] / p[2], 1 / p[2]])


def from_lambda(rep, v):
    if rep in (0, 2):
        return v.copy()
    if rep in (1, 3):
        return (1 / v[2]) * np.array([np.cos(v[0]) * np.sin(v[1]), np.sin(v[0]) * np.sin(v[1]), np.cos(v[1])])
    return np.array([v[0] / v[2], v[1] / v[2], 1 / v[2]])


def global_to_anchor(pose, calib, pG):
    R_GtoI, R_ItoC = jpl.quat_2_Rot(pose[:4]), jpl.quat_2_Rot(calib[:4])
    return R_ItoC @ (R_GtoI @ (pG - pose[4:7])) + calib[4:7]


def anchor_to_global(pose, calib, pA):
    R_GtoI, R_ItoC = jpl.quat_2_Rot(pose[:4]), jpl.quat_2_Rot(calib[:4])
    return R_GtoI.T @ (R_ItoC.T @ (pA - calib[4:7])) + pose[4:7]


def perturb(be, h, k, eps):
    """x <- x [+] eps e_k under the ov_type update rules (JPL left-multiplicative quaternion for poses)"""
    v, f = be.var_get(h)
    v = v.copy()
    if len(v) == 7:
        d = np.zeros(6)
        d[k] = eps
        dq = np.append(0.5 * d[:3], 1.0)
        v[:4] = jpl.quat_multiply(dq / np.linalg.norm(dq), v[:4])
        v[4:7] += d[3:]
    else:
        v[k] += eps
    be.var_set(h, v, f)


@pytest.mark.parametrize("anchor_in_track", [True, False])
@pytest.mark.parametrize("rep", [0, 1, 2, 3, 4])
def test_oracle_anchored_jacobian_vs_numerical_derivative(rep, anchor_in_track):
    S = synth.make_scenario("tiny_points", seed=1)
    S.options = dict(S.options, do_fej=0)
    orc = oracle_backend.OracleContext(S.options)
    ch = synth.load_scenario_into(orc, S)
    f = 3
    a, b = S.meas_offset[f], S.meas_offset[f + 1]
    track = [ch[i] for i in S.meas_clone_idx[a:b]]
    uv = S.uv[a:b]
    anchor = track[1] if anchor_in_track else [h for h in ch if h not in track][0] if len(track) < len(ch) else track[0]
    if not anchor_in_track and anchor in track:  # every clone observes this feature: drop one measurement to free a clone
        track, uv = track[:-1], uv[:-1]
        anchor = ch[S.meas_clone_idx[b - 1]]
    pG = S.p_FinG[f]
    calib = orc.var_get(orc.handle_calib())[0]
    pF = global_to_anchor(orc.var_get(anchor)[0], calib, pG) if rep >= 2 else pG
    lam = to_lambda(rep, pF)

    def resid(lam_):
        return orc.feature_jacobian_full_rep(track, uv, rep, anchor, from_lambda(rep, lam_), from_lambda(rep, lam_), 1.0)

    Hf, Hx, r0, order = resid(lam)
    assert Hf.shape == (2 * len(track), 3) and (anchor in order) == (rep >= 2 or anchor in track)
    eps = 1e-6
    for k in range(3):  # feature block: r = z - h(lambda) => dr/dlambda = -H_f
        d = np.zeros(3)
        d[k] = eps
        num = -(resid(lam + d)[2] - resid(lam - d)[2]) / (2 * eps)
        assert np.abs(num - Hf[:, k]).max() < 2e-6 * max(1.0, np.abs(Hf[:, k]).max()), (rep, k)
    col = 0
    for h in order:  # state blocks
        size = 6 if len(orc.var_get(h)[0]) == 7 else len(orc.var_get(h)[0])
        for k in range(size):
            v0, f0 = orc.var_get(h)
            perturb(orc, h, k, eps)
            rp = resid(lam)[2]
            orc.var_set(h, v0, f0)
            perturb(orc, h, k, -eps)
            rm = resid(lam)[2]
            orc.var_set(h, v0, f0)
            num = -(rp - rm) / (2 * eps)
            assert np.abs(num - Hx[:, col + k]).max() < 5e-6 * max(1.0, np.abs(Hx[:, col + k]).max()), (rep, h, k)
        col += size
    assert col == Hx.shape[1]


def _anchored_landmark(be, S, ch, rep, anchor):
    fid = int(S.slam[0][0])
    hl = be.slam_handle(fid)
    pG = be.var_get(hl)[0][:3].copy()
    calib = be.var_get(be.handle_calib())[0]
    pose_v, pose_f = be.var_get(anchor)
    be.var_set(hl, to_lambda(rep, global_to_anchor(pose_v, calib, pG)), to_lambda(rep, global_to_anchor(pose_f, calib, pG)))
    be.slam_set_representation(fid, rep, anchor)
    return fid, hl, pG


def _global_position_and_jacobian(be, fid, hl, order_handles, sizes):
    """p_FinG of an anchored landmark and d p_FinG / d [order_handles..., landmark] by central differences under the update rules"""
    rep, anchor = be.slam_get_representation(fid)

    def pG():
        calib = be.var_get(be.handle_calib())[0]
        return anchor_to_global(be.var_get(anchor)[0], calib, from_lambda(rep, be.var_get(hl)[0][:3]))

    p0, cols, eps = pG(), [], 1e-6
    for h, size in zip(list(order_handles) + [hl], list(sizes) + [3]):
        for k in range(size):
            v0, f0 = be.var_get(h)
            perturb(be, h, k, eps)
            pp = pG()
            be.var_set(h, v0, f0)
            perturb(be, h, k, -eps)
            pm = pG()
            be.var_set(h, v0, f0)
            cols.append((pp - pm) / (2 * eps))
    return p0, np.array(cols).T


@pytest.mark.parametrize("rep", [2, 3, 4])
def test_oracle_anchor_change_preserves_the_global_position_and_its_covariance(rep):
    S = synth.make_scenario("tiny_planes", seed=0)
    S.options = dict(S.options, do_fej=0)
    orc = oracle_backend.OracleContext(S.options)
    ch = synth.load_scenario_into(orc, S)
    fid, hl, pG0 = _anchored_landmark(orc, S, ch, rep, ch[2])
    hc = orc.handle_calib()
    ids = lambda hs, sz: np.concatenate([np.arange(orc.var_id(h), orc.var_id(h) + s) for h, s in zip(hs, sz)])
    hs, sz = [ch[2], hc, ch[7]], [6, 6, 6]
    sel = ids(hs + [hl], sz + [3])
    p_before, J_before = _global_position_and_jacobian(orc, fid, hl, hs, sz)
    P = orc.cov()
    C_before = J_before @ P[np.ix_(sel, sel)] @ J_before.T
    orc.slam_perform_anchor_change(fid, ch[7])
    assert orc.slam_get_representation(fid) == (rep, ch[7])
    p_after, J_after = _global_position_and_jacobian(orc, fid, hl, hs, sz)
    P2 = orc.cov()
    C_after = J_after @ P2[np.ix_(sel, sel)] @ J_after.T
    assert np.abs(p_after - pG0).max() < 1e-12 and np.abs(p_before - pG0).max() < 1e-12
    assert np.abs(C_after - C_before).max() < 1e-6 * np.abs(C_before).max(), (C_before, C_after)
    keep = np.setdiff1d(np.arange(P.shape[0]), np.arange(orc.var_id(hl), orc.var_id(hl) + 3))
    assert np.array_equal(P2[np.ix_(keep, keep)], P[np.ix_(keep, keep)])   # only the landmark's rows / columns change
    # cross-covariance of the global position with an unrelated variable (the IMU) is preserved as well
    him = np.arange(orc.var_id(orc.handle_imu()), orc.var_id(orc.handle_imu()) + 15)
    X_before, X_after = J_before @ P[np.ix_(sel, him)], J_after @ P2[np.ix_(sel, him)]
    assert np.abs(X_after - X_before).max() < 1e-6 * max(np.abs(X_before).max(), 1e-12)
