"""JPL-quaternion helpers (numpy) used by the synthetic-scenario generator and the host-side mirror.

Conventions follow ov_core `utils/quat_ops.h` as called by the reference (Propagator.cpp:384-404, UpdaterHelper.cpp:104,400):
quaternion `[x y z w]`, `R(q) = (2w^2-1) I - 2w [v x] + 2 v v^T`, left-multiplicative error `q <- dq (x) q`.
"""
import numpy as np


def skew(w):
    return np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])


def quat_2_Rot(q):
    v = np.asarray(q[:3], dtype=np.float64)
    w = float(q[3])
    return (2 * w * w - 1) * np.eye(3) - 2 * w * skew(v) + 2 * np.outer(v, v)


def rot_2_quat(rot):
    q = np.zeros(4)
    T = np.trace(rot)
    if rot[0, 0] >= T and rot[0, 0] >= rot[1, 1] and rot[0, 0] >= rot[2, 2]:
        q[0] = np.sqrt((1 + 2 * rot[0, 0] - T) / 4)
        q[1] = (1 / (4 * q[0])) * (rot[0, 1] + rot[1, 0])
        q[2] = (1 / (4 * q[0])) * (rot[0, 2] + rot[2, 0])
        q[3] = (1 / (4 * q[0])) * (rot[1, 2] - rot[2, 1])
    elif rot[1, 1] >= T and rot[1, 1] >= rot[0, 0] and rot[1, 1] >= rot[2, 2]:
        q[1] = np.sqrt((1 + 2 * rot[1, 1] - T) / 4)
        q[0] = (1 / (4 * q[1])) * (rot[0, 1] + rot[1, 0])
        q[2] = (1 / (4 * q[1])) * (rot[1, 2] + rot[2, 1])
        q[3] = (1 / (4 * q[1])) * (rot[2, 0] - rot[0, 2])
    elif rot[2, 2] >= T and rot[2, 2] >= rot[0, 0] and rot[2, 2] >= rot[1, 1]:
        q[2] = np.sqrt((1 + 2 * rot[2, 2] - T) / 4)
        q[0] = (1 / (4 * q[2])) * (rot[0, 2] + rot[2, 0])
        q[1] = (1 / (4 * q[2])) * (rot[1, 2] + rot[2, 1])
        q[3] = (1 / (4 * q[2])) * (rot[0, 1] - rot[1, 0])
    else:
        q[3] = np.sqrt((1 + T) / 4)
        q[0] = (1 / (4 * q[3])) * (rot[1, 2] - rot[2, 1])
        q[1] = (1 / (4 * q[3])) * (rot[2, 0] - rot[0, 2])
        q[2] = (1 / (4 * q[3])) * (rot[0, 1] - rot[1, 0])
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def quat_multiply(q, p):
    Qm = np.zeros((4, 4))
    Qm[:3, :3] = q[3] * np.eye(3) - skew(q[:3])
    Qm[:3, 3] = q[:3]
    Qm[3, :3] = -q[:3]
    Qm[3, 3] = q[3]
    r = Qm @ p
    if r[3] < 0:
        r = -r
    return r / np.linalg.norm(r)


def quat_left_update(q, dtheta):
    dq = np.array([0.5 * dtheta[0], 0.5 * dtheta[1], 0.5 * dtheta[2], 1.0])
    dq = dq / np.linalg.norm(dq)
    return quat_multiply(dq, q)


def exp_so3(w):
    th = np.linalg.norm(w)
    if th == 0:
        return np.eye(3)
    wx = skew(w)
    if th < 1e-7:
        A, B = 1.0, 0.5
    else:
        A, B = np.sin(th) / th, (1 - np.cos(th)) / th ** 2
    return np.eye(3) + A * wx + B * wx @ wx


def radtan_distort(cam, xn, yn):
    fx, fy, cx, cy, k1, k2, p1, p2 = cam
    r2 = xn * xn + yn * yn
    r4 = r2 * r2
    x1 = xn * (1 + k1 * r2 + k2 * r4) + 2 * p1 * xn * yn + p2 * (r2 + 2 * xn * xn)
    y1 = yn * (1 + k1 * r2 + k2 * r4) + p1 * (r2 + 2 * yn * yn) + 2 * p2 * xn * yn
    return fx * x1 + cx, fy * y1 + cy
