"""The fused Cholesky kernel (csrc/cholfused.cu) on its own against NumPy: sizes that are not multiples of the 64 / 16 blocking,
partial factorisations (rows below the pivoted columns are solved along), the zero-pivot rule on rank-deficient Gram matrices, and
the fused triangular solve Y = M L^-T, w = L^-1 z."""
import ctypes as C

import numpy as np
import pytest

from ov_plane_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    S = synth.make_scenario("tiny_points")
    c = api.Context(S.options, device=0, max_state=640, max_meas_rows=4096, debug=True)  # libovp_debug.so: ovp_debug_chol_solve (include/ovp_debug.h)
    yield c
    c.close()


def _run(ctx, A, npiv, tol, M=None, z=None):
    n = A.shape[0]
    A = np.asfortranarray(np.tril(A))
    L = np.zeros((n, n), order="F")
    Y = w = None
    pM = pz = pY = pw = None
    mrows = 0
    if M is not None:
        M = np.asfortranarray(M)
        z = np.ascontiguousarray(z)
        mrows = M.shape[0]
        Y, w = np.zeros((mrows, npiv), order="F"), np.zeros(npiv)
        pM, pz, pY, pw = (x.ctypes.data_as(C.c_void_p) for x in (M, z, Y, w))
    ctx._ck(ctx.lib.ovp_debug_chol_solve(ctx.h, A.ctypes.data_as(C.c_void_p), n, npiv, C.c_double(tol), pM, mrows, pz,
                                         L.ctypes.data_as(C.c_void_p), pY, pw))
    return np.tril(L), Y, w


@pytest.mark.parametrize("n", [1, 5, 16, 17, 63, 64, 65, 100, 129, 200, 333, 474, 600])
def test_spd_factor_and_solve(ctx, n):
    rng = np.random.default_rng(n)
    B = rng.normal(size=(n + 8, n))
    A = B.T @ B + 0.5 * np.eye(n)
    M = rng.normal(size=(37 + n % 29, n))
    z = rng.normal(size=n)
    L, Y, w = _run(ctx, A, n, 0.0, M, z)
    Lr = np.linalg.cholesky(A)
    sc = np.abs(Lr).max()
    assert np.abs(L - Lr).max() < 1e-11 * sc * n
    assert np.abs(Y - np.linalg.solve(Lr, M.T).T).max() < 1e-9 * np.abs(M).max() * n
    assert np.abs(w - np.linalg.solve(Lr, z)).max() < 1e-9 * np.abs(z).max() * n


@pytest.mark.parametrize("n,npiv", [(70, 64), (138, 134), (200, 130), (471, 470), (300, 17)])
def test_partial_factorisation_solves_the_rows_below(ctx, n, npiv):
    rng = np.random.default_rng(n + npiv)
    B = rng.normal(size=(n + 5, n))
    A = B.T @ B + np.eye(n)
    L, _, _ = _run(ctx, A, npiv, 0.0)
    L11 = np.linalg.cholesky(A[:npiv, :npiv])
    L21 = np.linalg.solve(L11, A[npiv:, :npiv].T).T
    assert np.abs(L[:npiv, :npiv] - L11).max() < 1e-11 * np.abs(L11).max() * n
    assert np.abs(L[npiv:, :npiv] - L21).max() < 1e-10 * np.abs(L21).max() * n


def test_zero_pivot_rule_on_a_rank_deficient_gram_matrix(ctx):
    rng = np.random.default_rng(7)
    n, r = 150, 120
    H = rng.normal(size=(400, r)) @ rng.normal(size=(r, n))  # rank r < n
    G = H.T @ H
    L, _, _ = _run(ctx, G, n, 1e-11)
    assert np.isfinite(L).all()
    zero_cols = np.where(np.abs(np.diag(L)) == 0.0)[0]
    assert len(zero_cols) == n - r                      # exactly the rank deficiency is rejected
    assert np.abs(L[:, zero_cols]).max() == 0.0         # rejected pivots leave a column of zeros
    assert np.abs(L @ L.T - G).max() < 1e-8 * np.abs(G).max()   # and L L^T still reproduces G


def test_strict_mode_reports_an_indefinite_matrix(ctx):
    A = np.eye(80)
    A[40, 40] = -1.0
    with pytest.raises(api.OvpError):
        _run(ctx, A, 80, 0.0)


def test_two_launch_fallback_when_the_grid_cannot_be_co_resident():
    """n = 1000 with 700 right-hand-side rows needs 136 tile CTAs + 44 row-block CTAs > 148 SMs: the factorisation and the triangular solve
    run as two launches (the second reads the factor tiles from the exchange slots without flags)."""
    S = synth.make_scenario("tiny_points")
    c = api.Context(S.options, device=0, max_state=1024, max_meas_rows=4096, debug=True)
    rng = np.random.default_rng(5)
    n = 1000
    B = rng.normal(size=(n + 16, n))
    A = B.T @ B + 0.5 * np.eye(n)
    M = rng.normal(size=(700, n))
    z = rng.normal(size=n)
    L, Y, w = _run(c, A, n, 0.0, M, z)
    Lr = np.linalg.cholesky(A)
    assert np.abs(L - Lr).max() < 1e-11 * np.abs(Lr).max() * n
    assert np.abs(Y - np.linalg.solve(Lr, M.T).T).max() < 1e-9 * np.abs(M).max() * n
    assert np.abs(w - np.linalg.solve(Lr, z)).max() < 1e-9 * np.abs(z).max() * n
    c.close()
