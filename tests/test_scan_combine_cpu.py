"""CPU pins of the scan's filtering / smoothing combine (host build of the operator in scan.cu,
exported as bdlm_scan_combine -- the same source the kernels inline).

The kernels form BOTH factors of the filtering combine from ONE inverse, using
(I + J_j C_i) = (I + C_i J_j)^T for symmetric C_i, J_j.  These tests hold that against

* an independent numpy restatement with the two separate solves of the published operator
  (Sarkka & Garcia-Fernandez 2021, eq. for a_i (x) a_j), and
* the reference recursion itself: folding the per-observation elements left to right must give
  the oracle's Kalman filter (KalmanFilter.scala:64-118 via oracle.kf_filter) and folding the
  smoothing elements right to left the textbook RTS smoother, missing observations included.
"""
import numpy as np
import pytest

import oracle
from bayesian_dlms_b200 import _capi as capi, dlm
import helpers as H


def _cm(a):  # column-major flattening, as the ABI stores matrices
    return np.asarray(a, float).T.ravel()


def _pack_f(A, b, Cm, eta, J):
    return np.concatenate([_cm(A), b, _cm(Cm), eta, _cm(J)])


def _unpack_f(x, n):
    nn = n * n
    A = x[:nn].reshape(n, n).T
    b = x[nn:nn + n]
    Cm = x[nn + n:2 * nn + n].reshape(n, n).T
    eta = x[2 * nn + n:2 * nn + 2 * n]
    J = x[2 * nn + 2 * n:].reshape(n, n).T
    return A, b, Cm, eta, J


def _comb(lib, n, backward, x, y):
    out = np.empty_like(x)
    assert lib.bdlm_scan_combine(n, backward, x.ctypes.data, y.ctypes.data, out.ctypes.data) == 0
    return out


def _np_f_combine(ei, ej):
    """a_i (x) a_j with the two solves written out (i earlier in time)."""
    Ai, bi, Ci, etai, Ji = ei
    Aj, bj, Cj, etaj, Jj = ej
    n = bi.size
    X = Aj @ np.linalg.inv(np.eye(n) + Ci @ Jj)
    Y = Ai.T @ np.linalg.inv(np.eye(n) + Jj @ Ci)
    return (X @ Ai, X @ (bi + Ci @ etaj) + bj, X @ Ci @ Aj.T + Cj,
            Y @ (etaj - Jj @ bi) + etai, Y @ Jj @ Ai + Ji)


@pytest.mark.parametrize("n", [1, 2, 3, 4])
def test_single_inverse_combine_equals_the_two_solve_operator(n):
    lib = capi.load()
    rng = np.random.default_rng(100 + n)

    def spd(scale):
        a = rng.standard_normal((n, n))
        return scale * (a @ a.T / n + 0.2 * np.eye(n))

    for _ in range(20):
        ei = (rng.standard_normal((n, n)) * 0.7, rng.standard_normal(n), spd(3.0), rng.standard_normal(n), spd(0.5))
        ej = (rng.standard_normal((n, n)) * 0.7, rng.standard_normal(n), spd(2.0), rng.standard_normal(n), spd(0.8))
        got = _unpack_f(_comb(lib, n, 0, _pack_f(*ei), _pack_f(*ej)), n)
        want = _np_f_combine(ei, ej)
        for g, w, name in zip(got, want, "A b C eta J".split()):
            assert np.allclose(g, w, rtol=1e-10, atol=1e-12), (name, np.abs(g - w).max())


def _f_element(G, F, W, V, y):
    """Filtering element of one observation (p = 1, unit step: Q = W); a missing observation is
    a pure prediction."""
    n = G.shape[0]
    if np.isnan(y):
        return G.copy(), np.zeros(n), W.copy(), np.zeros(n), np.zeros((n, n))
    S = float(F @ W @ F + V)
    K = W @ F / S
    IKH = np.eye(n) - np.outer(K, F)
    GtF = G.T @ F
    return IKH @ G, K * y, IKH @ W, GtF * y / S, np.outer(GtF, GtF) / S


@pytest.mark.parametrize("n", [1, 2, 3])
def test_folding_elements_reproduces_the_reference_filter_and_textbook_smoother(n):
    lib = capi.load()
    rng = np.random.default_rng(7 + n)
    mod = dlm.polynomial(n)
    T = 60
    V = np.array([[1.7]])
    W = np.diag(np.linspace(0.9, 0.3, n))
    m0, C0 = np.zeros(n), 4.0 * np.eye(n)
    times = np.arange(1, T + 1.0)
    y = H.simulate(mod, V, W, m0, C0, times, rng, missing=0.15)[:, 0]
    G, F = mod.g(1.0), mod.f(1.0)[:, 0]
    kf = oracle.kf_filter(n, 1, _cm(mod.f(1.0)), _cm(G), _cm(V), _cm(W), m0, _cm(C0), times,
                          y.reshape(-1, 1), keep_init=True)   # the oracle takes column-major matrices

    # forward: state element (A = 0, b = m0, C = C0) folded with one element per observation
    acc = _pack_f(np.zeros((n, n)), m0, C0, np.zeros(n), np.zeros((n, n)))
    for t in range(T):
        acc = _comb(lib, n, 0, acc, _pack_f(*_f_element(G, F, W, float(V[0, 0]), y[t])))
        A, b, Cm, _, _ = _unpack_f(acc, n)
        assert np.abs(A).max() == 0.0                      # a prefix that starts at a state stays a state
        assert np.allclose(b, kf["m"][t + 1], rtol=1e-10, atol=1e-12), t
        assert np.allclose(_cm(Cm), kf["C"][t + 1], rtol=1e-10, atol=1e-12), t

    # backward: terminal element (E = 0, g = m_T, L = C_T), elements E = C G^T R1^-1,
    # g = m - E a1, L = C - E R1 E^T folded from the right = the textbook RTS recursion
    m = kf["m"]
    Cs = kf["C"].reshape(T + 1, n, n).transpose(0, 2, 1)
    s, S = m[T].copy(), Cs[T].copy()
    acc = np.concatenate([np.zeros(n * n), s, _cm(S)])
    for r in range(T - 1, -1, -1):
        a1 = G @ m[r]
        R1 = G @ Cs[r] @ G.T + W
        E = Cs[r] @ G.T @ np.linalg.inv(R1)
        s = m[r] + E @ (s - a1)                              # Smoothing.scala:31-47 with the transpose in place
        S = Cs[r] - E @ (R1 - S) @ E.T
        elem = np.concatenate([_cm(E), m[r] - E @ a1, _cm(Cs[r] - E @ R1 @ E.T)])
        acc = _comb(lib, n, 1, elem, acc)
        assert np.allclose(acc[n * n:n * n + n], s, rtol=1e-9, atol=1e-11), r
        assert np.allclose(acc[n * n + n:], _cm(S), rtol=1e-9, atol=1e-11), r
