"""CPU pin of the two-lanes-per-series arithmetic (csrc/pair_steps.cuh, kernel in csrc/kf_pair.cu).

The n = 4 kernel splits every 4 x 4 matrix of a series over two lanes and gathers halves between
them.  bdlm_debug_pair_filter_smooth_host is a HOST build of those very step functions, with two
host threads standing in for the two lanes, so the split operation order can be held against the
oracle (KalmanFilter.scala:64-118,273-321; Smoothing.scala:31-64 via oracle/bdlm_oracle.c)
without a GPU: every output must be bit-identical, in both smoother modes, with missing
observations, on irregular grids and with repeated time stamps (dt == 0).
"""
import numpy as np
import pytest

import oracle
from bayesian_dlms_b200 import _capi as capi, dlm
import helpers as H

N = 4


def _cm(a):
    return np.ascontiguousarray(np.asarray(a, float).T).ravel()


def _host_pair(G, F, V, W, m0, C0, dt, y, textbook):
    lib = capi.load()
    T = y.size
    out = {k: np.full((T + 1, d), -7.0) for k, d in
           dict(m=N, C=N * N, a=N, R=N * N, f=1, Q=1, s=N, S=N * N).items()}
    Gc, Wc, C0c = _cm(G), _cm(W), _cm(C0)
    Fv, m0v, yv = np.ascontiguousarray(F, float), np.ascontiguousarray(m0, float), np.ascontiguousarray(y, float)
    dtv = None if dt is None else np.ascontiguousarray(dt, float)
    st = lib.bdlm_debug_pair_filter_smooth_host(
        Gc.ctypes.data, Fv.ctypes.data, float(V), Wc.ctypes.data, m0v.ctypes.data, C0c.ctypes.data,
        None if dtv is None else dtv.ctypes.data, yv.ctypes.data, T, int(textbook),
        *(out[k].ctypes.data for k in ("m", "C", "a", "R", "f", "Q", "s", "S")))
    out["status"] = st
    return out


def _oracle(G, F, V, W, m0, C0, times, y, textbook):
    kf = oracle.kf_filter(N, 1, np.asarray(F, float).reshape(N, 1), _cm(G), [float(V)], _cm(W), m0, _cm(C0),
                          times, y.reshape(-1, 1), keep_init=True)
    sm = oracle.rts_smooth(N, _cm(G), kf, keep_init=True, textbook=textbook)
    kf.update(s=sm["s"], S=sm["S"], status=kf["status"] | sm["status"])
    return kf


def _spd(rng, scale):
    a = rng.standard_normal((N, N))
    return scale * (a @ a.T / N + 0.3 * np.eye(N))


def _check(got, want):
    for k in ("m", "C", "a", "R", "f", "Q", "s", "S"):
        g, w = got[k], want[k].reshape(got[k].shape)
        assert np.array_equal(g, w, equal_nan=True), (k, np.nanmax(np.abs(g - w)))
    assert got["status"] == want["status"]


@pytest.mark.parametrize("textbook", [False, True])
@pytest.mark.parametrize("model", ["polynomial4", "dense"])
def test_pair_arithmetic_is_bit_identical_to_the_oracle_regular_grid(model, textbook):
    rng = np.random.default_rng(11 + (model == "dense") + 2 * textbook)
    if model == "polynomial4":
        mod = dlm.polynomial(4)
        G, F = mod.g(1.0), mod.f(1.0)[:, 0]
        W = np.diag([1.0, 0.5, 0.1, 0.05])
    else:  # no structural zeros anywhere: every product term is exercised
        G = 0.6 * rng.standard_normal((N, N)) + 0.5 * np.eye(N)
        F = rng.standard_normal(N)
        W = _spd(rng, 0.7)
    V, m0, C0 = 1.5, rng.standard_normal(N), _spd(rng, 5.0)
    T = 90
    times = np.arange(1, T + 1.0)
    y = np.cumsum(rng.standard_normal(T)) + rng.standard_normal(T)
    y[rng.random(T) < 0.15] = np.nan
    got = _host_pair(G, F, V, W, m0, C0, None, y, textbook)
    want = _oracle(G, F, V, W, m0, C0, times, y, textbook)
    _check(got, want)
    assert got["status"] == 0


@pytest.mark.parametrize("textbook", [False, True])
def test_pair_arithmetic_irregular_grid_with_repeated_time_stamps(textbook):
    """dt != 1 scales W, dt == 0 passes the state through (KalmanFilter.scala:273-286) -- also
    twice in a row with the observation missing, the path on which two exchanges through the
    same buffer follow each other."""
    rng = np.random.default_rng(5 + textbook)
    G = 0.5 * rng.standard_normal((N, N)) + 0.6 * np.eye(N)
    F = rng.standard_normal(N)
    W, V, m0, C0 = _spd(rng, 0.4), 0.8, rng.standard_normal(N), _spd(rng, 3.0)
    T = 70
    dt = rng.choice([0.0, 0.5, 1.0, 1.0, 2.0, 3.25], size=T)
    dt[0] = 1.0                      # t0 = min(time) - 1
    dt[10:13] = 0.0
    times = np.cumsum(dt)
    y = rng.standard_normal(T) * 2.0
    y[rng.random(T) < 0.2] = np.nan
    y[10:13] = np.nan
    got = _host_pair(G, F, V, W, m0, C0, dt, y, textbook)
    want = _oracle(G, F, V, W, m0, C0, times, y, textbook)
    _check(got, want)


def test_pair_arithmetic_flags_a_singular_step_like_the_oracle():
    """W = 0, C0 = 0, V = 0: Q = 0 at the first observed step and R is singular in the smoother;
    the status bits (not the garbage values) must agree with the oracle."""
    G, F = np.eye(N), np.array([1.0, 0.0, 0.0, 0.0])
    W, C0, m0 = np.zeros((N, N)), np.zeros((N, N)), np.zeros(N)
    T = 6
    y = np.arange(1.0, T + 1.0)
    got = _host_pair(G, F, 0.0, W, m0, C0, None, y, False)
    want = _oracle(G, F, 0.0, W, m0, C0, np.arange(1, T + 1.0), y, False)
    # the kernels add BDLM_ST_NONFINITE for a NaN / Inf final state; the oracle reports the pivots only
    assert want["status"] == capi.ST_SINGULAR
    assert got["status"] & ~capi.ST_NONFINITE == want["status"]
    assert (got["status"] & capi.ST_NONFINITE != 0) == (not np.isfinite(want["s"][0]).all())


def test_debug_pair_mode_validates_its_argument():
    lib = capi.load()
    assert lib.bdlm_debug_set_pair_mode(7) == capi.E_ARG
    assert lib.bdlm_debug_set_pair_mode(0) == 0
