"""ctypes driver for oracle/bdlm_oracle.c (the CPU restatement of the reference).

TEST INFRASTRUCTURE ONLY: see the header of bdlm_oracle.c.  All arrays are
series-major ``[rows][k]`` with column-major matrices inside a row (Breeze
``DenseMatrix.data`` order), fp64, NaN standing for a missing observation
(``None`` in the reference's ``DenseVector[Option[Double]]``, Dlm.scala:94).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "bdlm_oracle.c")
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None

_dp = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    """gcc -O2, no FMA contraction (the JVM never fuses), OpenMP for the batch drivers."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC",
               "-fvisibility=hidden", "-o", _SO, _SRC, "-lm"]
        subprocess.run(cmd, check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _a(x, shape=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    if shape is not None:
        x = x.reshape(shape)
    return x


def _p(x):
    return x.ctypes.data_as(_dp)


def _model(F, G, n, p, T):
    """F: (n,p) column-major flattened as (n*p,) or (T, n*p); same for G."""
    F = _a(F)
    G = _a(G)
    f_tv = int(F.size == T * n * p and T * n * p != n * p)
    g_tv = int(G.size == T * n * n and T * n * n != n * n)
    assert F.size == (T if f_tv else 1) * n * p, (F.shape, n, p, T)
    assert G.size == (T if g_tv else 1) * n * n, (G.shape, n, T)
    return F, f_tv, G, g_tv


def cm(M):
    """numpy (r,c) matrix -> column-major flat (Breeze order)."""
    return np.ascontiguousarray(np.asarray(M, dtype=np.float64).T).ravel()


def kf_filter(n, p, F, G, V, W, m0, C0, times, y, keep_init=True, v_tv=False, t_init=None,
              w_tv=False):
    """v_tv / w_tv: V / W hold T matrices (StudentTGibbs.filter, StudentTGibbs.scala:100-119;
    DlmFsvSystem.ffbs, DlmFsvSystem.scala:137-167)."""
    times = _a(times)
    T = times.size
    if v_tv:
        assert _a(V).size == T * p * p
    if w_tv:
        assert _a(W).size == T * n * n
    y = _a(y, (T, p))
    F, f_tv, G, g_tv = _model(F, G, n, p, T)
    rows = T + int(keep_init)
    out = {k: np.empty((rows, d)) for k, d in
           dict(m=n, C=n * n, a=n, R=n * n, f=p, Q=p * p).items()}
    tm = np.empty(rows)
    if t_init is not None:
        assert not v_tv
        st = lib().oracle_kf_filter_from(
            n, p, T, _p(F), f_tv, _p(G), g_tv, _p(_a(V)), _p(_a(W)), _p(_a(m0)), _p(_a(C0)),
            C.c_double(float(t_init)), _p(times), _p(y), int(keep_init), _p(tm),
            *(_p(out[k]) for k in ("m", "C", "a", "R", "f", "Q")))
        out["time"] = tm
        out["status"] = st
        return out
    st = lib().oracle_kf_filter_tv(
        n, p, T, _p(F), f_tv, _p(G), g_tv, _p(_a(V)), int(v_tv), _p(_a(W)), int(w_tv),
        _p(_a(m0)), _p(_a(C0)), _p(times), _p(y), int(keep_init), _p(tm),
        *(_p(out[k]) for k in ("m", "C", "a", "R", "f", "Q")))
    out["time"] = tm
    out["status"] = st
    return out


def rts_smooth(n, G, filt, keep_init=True, textbook=False):
    rows = filt["m"].shape[0]
    T = rows - int(keep_init)
    G = _a(G)
    g_tv = int(G.size == T * n * n and T > 1)
    s = np.empty((rows, n))
    S = np.empty((rows, n * n))
    st = lib().oracle_rts_smooth(n, T, int(keep_init), _p(G), g_tv, _p(filt["m"]),
                                 _p(filt["C"]), _p(filt["a"]), _p(filt["R"]),
                                 int(textbook), _p(s), _p(S))
    return dict(s=s, S=S, status=st)


def backward_sample(n, G, W, filt, z, keep_init=True):
    rows = filt["m"].shape[0]
    T = rows - int(keep_init)
    G = _a(G)
    g_tv = int(G.size == T * n * n and T > 1)
    z = _a(z, (rows, n))
    theta = np.empty((rows, n))
    st = lib().oracle_backward_sample(n, T, int(keep_init), _p(G), g_tv, _p(_a(W)),
                                      _p(filt["time"]), _p(filt["m"]), _p(filt["C"]),
                                      _p(filt["a"]), _p(filt["R"]), _p(z), _p(theta))
    return dict(theta=theta, status=st)


def ffbs(n, p, F, G, V, W, m0, C0, times, y, z, v_tv=False, w_tv=False):
    times = _a(times)
    T = times.size
    rows = T + 1
    y = _a(y, (T, p))
    z = _a(z, (rows, n))
    F, f_tv, G, g_tv = _model(F, G, n, p, T)
    out = {k: np.empty((rows, d)) for k, d in
           dict(theta=n, m=n, C=n * n, a=n, R=n * n).items()}
    tm = np.empty(rows)
    st = lib().oracle_ffbs_tv(n, p, T, _p(F), f_tv, _p(G), g_tv, _p(_a(V)), int(v_tv), _p(_a(W)), int(w_tv),
                           _p(_a(m0)), _p(_a(C0)), _p(times), _p(y), _p(z), _p(tm),
                           *(_p(out[k]) for k in ("theta", "m", "C", "a", "R")))
    out["time"] = tm
    out["status"] = st
    return out


def loglik(n, p, F, G, V, W, m0, C0, times, y):
    times = _a(times)
    T = times.size
    y = _a(y, (T, p))
    F, f_tv, G, g_tv = _model(F, G, n, p, T)
    tr = C.c_double()
    inn = C.c_double()
    st = lib().oracle_loglik(n, p, T, _p(F), f_tv, _p(G), g_tv, _p(_a(V)), _p(_a(W)),
                             _p(_a(m0)), _p(_a(C0)), _p(times), _p(y),
                             C.byref(tr), C.byref(inn))
    return dict(transition=tr.value, innovations=inn.value, status=st)


def sqrt_svd(M, inv=False):
    M = _a(M)
    n = int(round(M.size ** 0.5))
    out = np.empty(n * n)
    lib().oracle_sqrt_svd(n, _p(M), int(inv), _p(out))
    return out


def svd_filter(n, p, F, G, V, Wadv, m0, C0, times, y, keep_init=True, transform=True):
    times = _a(times)
    T = times.size
    y = _a(y, (T, p))
    F, f_tv, G, g_tv = _model(F, G, n, p, T)
    rows = T + int(keep_init)
    out = {k: np.empty((rows, d)) for k, d in
           dict(m=n, dc=n, uc=n * n, a=n, dr=n, ur=n * n, f=p).items()}
    tm = np.empty(rows)
    st = lib().oracle_svd_filter(
        n, p, T, _p(F), f_tv, _p(G), g_tv, _p(_a(V)), int(transform), _p(_a(Wadv)),
        _p(_a(m0)), _p(_a(C0)), _p(times), _p(y), int(keep_init), _p(tm),
        *(_p(out[k]) for k in ("m", "dc", "uc", "a", "dr", "ur", "f")))
    out["time"] = tm
    out["status"] = st
    return out


def svd_filter_tv(n, p, F, G, V, W, m0, C0, times, y, keep_init=True, v_tv=False, w_tv=False,
                  consistent=True):
    """Next row f2 on the SVD path: V [T][p*p] / W [T][n*n] raw matrices per observation
    (DlmFsv.ffbsSvd, DlmFsv.scala:208-229; DlmFsvSystem.ffbsSvd, DlmFsvSystem.scala:177-207)."""
    times = _a(times)
    T = times.size
    y = _a(y, (T, p))
    F, f_tv, G, g_tv = _model(F, G, n, p, T)
    assert _a(V).size == (T if v_tv else 1) * p * p and _a(W).size == (T if w_tv else 1) * n * n
    rows = T + int(keep_init)
    out = {k: np.empty((rows, d)) for k, d in
           dict(m=n, dc=n, uc=n * n, a=n, dr=n, ur=n * n, f=p).items()}
    tm = np.empty(rows)
    st = lib().oracle_svd_filter_tv(
        n, p, T, _p(F), f_tv, _p(G), g_tv, _p(_a(V)), int(v_tv), _p(_a(W)), int(w_tv),
        int(consistent), _p(_a(m0)), _p(_a(C0)), _p(times), _p(y), int(keep_init), _p(tm),
        *(_p(out[k]) for k in ("m", "dc", "uc", "a", "dr", "ur", "f")))
    out["time"] = tm
    out["status"] = st
    return out


def svd_ffbs_tv(n, p, F, G, V, W, m0, C0, times, y, z, v_tv=False, w_tv=False, consistent=True):
    times = _a(times)
    T = times.size
    rows = T + 1
    y = _a(y, (T, p))
    z = _a(z, (rows, n))
    F, f_tv, G, g_tv = _model(F, G, n, p, T)
    assert _a(V).size == (T if v_tv else 1) * p * p and _a(W).size == (T if w_tv else 1) * n * n
    out = {k: np.empty((rows, d)) for k, d in
           dict(theta=n, m=n, dc=n, uc=n * n, a=n, dr=n, ur=n * n).items()}
    tm = np.empty(rows)
    st = lib().oracle_svd_ffbs_tv(
        n, p, T, _p(F), f_tv, _p(G), g_tv, _p(_a(V)), int(v_tv), _p(_a(W)), int(w_tv),
        int(consistent), _p(_a(m0)), _p(_a(C0)), _p(times), _p(y), _p(z), _p(tm),
        *(_p(out[k]) for k in ("theta", "m", "dc", "uc", "a", "dr", "ur")))
    out["time"] = tm
    out["status"] = st
    return out


def svd_backward_sample(n, G, sqrtW, filt, z, keep_init=True):
    rows = filt["m"].shape[0]
    T = rows - int(keep_init)
    G = _a(G)
    g_tv = int(G.size == T * n * n and T > 1)
    z = _a(z, (rows, n))
    theta = np.empty((rows, n))
    st = lib().oracle_svd_backward_sample(n, T, int(keep_init), _p(G), g_tv,
                                          _p(_a(sqrtW)), _p(filt["m"]), _p(filt["dc"]),
                                          _p(filt["uc"]), _p(filt["a"]), _p(z), _p(theta))
    return dict(theta=theta, status=st)


def svd_ffbs(n, p, F, G, V, W, m0, C0, times, y, z, consistent=False):
    times = _a(times)
    T = times.size
    rows = T + 1
    y = _a(y, (T, p))
    z = _a(z, (rows, n))
    F, f_tv, G, g_tv = _model(F, G, n, p, T)
    out = {k: np.empty((rows, d)) for k, d in
           dict(theta=n, m=n, dc=n, uc=n * n, a=n, dr=n, ur=n * n).items()}
    tm = np.empty(rows)
    st = lib().oracle_svd_ffbs(n, p, T, _p(F), f_tv, _p(G), g_tv, _p(_a(V)), _p(_a(W)),
                               _p(_a(m0)), _p(_a(C0)), _p(times), _p(y), _p(z),
                               int(consistent), _p(tm),
                               *(_p(out[k]) for k in
                                 ("theta", "m", "dc", "uc", "a", "dr", "ur")))
    out["time"] = tm
    out["status"] = st
    return out


def gibbs_stats(n, p, F, G, times_rows, y, theta):
    times_rows = _a(times_rows)
    T = times_rows.size - 1
    y = _a(y, (T, p))
    theta = _a(theta, (T + 1, n))
    F, f_tv, G, g_tv = _model(F, G, n, p, T)
    ssy, ny, ssw, sc = np.empty(p), np.empty(p), np.empty(n), np.empty(n * n)
    lib().oracle_gibbs_stats.restype = None
    lib().oracle_gibbs_stats(n, p, T, _p(F), f_tv, _p(G), g_tv, _p(times_rows), _p(y),
                             _p(theta), _p(ssy), _p(ny), _p(ssw), _p(sc))
    return dict(ssy=ssy, ny=ny, ssw=ssw, scatter=sc)


def eigsym(A):
    A = _a(A)
    n = int(round(A.size ** 0.5))
    lam, V = np.empty(n), np.empty(n * n)
    st = lib().oracle_eigsym(n, _p(A), _p(lam), _p(V))
    return lam, V.reshape(n, n).T.copy(), st  # V as numpy (row, col)


def svd(M, r, n):
    M = _a(M)
    sv, V = np.empty(n), np.empty(n * n)
    st = lib().oracle_svd(r, n, _p(M), _p(sv), _p(V))
    return sv, V.reshape(n, n).T.copy(), st


def solve(A, B):
    """A (n,n) numpy, B (n,nrhs) numpy -> X numpy via the dgesv restatement."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    n, nrhs = B.shape
    X = np.empty(n * nrhs)
    st = lib().oracle_solve(n, _p(cm(A)), nrhs, _p(cm(B)), _p(X))
    return X.reshape(nrhs, n).T.copy(), st


def _stride(x, k, B):
    x = _a(x)
    return x, (0 if x.size == k else k)


def batch_filter_smooth(B, n, p, T, F, G, V, W, m0, C0, times, y, keep_init=True,
                        nthreads=None, out=None):
    """[B][rows][k] outputs; V/W/m0/C0 either shared (k values) or per series (B*k).
    `out` lets a timing loop reuse the output arrays of a previous call."""
    nthreads = nthreads or os.cpu_count()
    rows = T + int(keep_init)
    V, vs = _stride(V, p * p, B)
    W, ws = _stride(W, n * n, B)
    m0, ms = _stride(m0, n, B)
    C0, cs = _stride(C0, n * n, B)
    y = _a(y, (B, T, p))
    if out is None:
        out = {k: np.empty((B, rows, d)) for k, d in
               dict(m=n, C=n * n, a=n, R=n * n, f=p, Q=p * p, s=n, S=n * n).items()}
    L = lib()
    L.oracle_batch_filter_smooth.argtypes = (
        [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp] +
        [_dp, C.c_int64] * 4 + [_dp, _dp, C.c_int] + [_dp] * 8)
    bad = L.oracle_batch_filter_smooth(
        B, nthreads, n, p, T, _p(_a(F)), _p(_a(G)), _p(V), vs, _p(W), ws, _p(m0), ms,
        _p(C0), cs, _p(_a(times)), _p(y), int(keep_init),
        *(_p(out[k]) for k in ("m", "C", "a", "R", "f", "Q", "s", "S")))
    out["bad"] = bad
    return out


def batch_ffbs(B, n, p, T, F, G, V, W, m0, C0, times, y, z, svd=False, nthreads=None):
    nthreads = nthreads or os.cpu_count()
    rows = T + 1
    V, vs = _stride(V, p * p, B)
    W, ws = _stride(W, n * n, B)
    m0, ms = _stride(m0, n, B)
    C0, cs = _stride(C0, n * n, B)
    y = _a(y, (B, T, p))
    z = _a(z, (B, rows, n))
    theta = np.empty((B, rows, n))
    L = lib()
    L.oracle_batch_ffbs.argtypes = (
        [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp] +
        [_dp, C.c_int64] * 4 + [_dp, _dp, _dp, C.c_int, _dp])
    bad = L.oracle_batch_ffbs(B, nthreads, n, p, T, _p(_a(F)), _p(_a(G)), _p(V), vs,
                              _p(W), ws, _p(m0), ms, _p(C0), cs, _p(_a(times)), _p(y),
                              _p(z), int(svd), _p(theta))
    return dict(theta=theta, bad=bad)


# ---- "next" rows: conjugate draws, AR(1)/OU scalar filters, conjugate filter

def gibbs_invgamma(prior_shape, prior_scale, ss, g, count=None, count_all=0.0):
    """Posterior InverseGamma(shape, rate) per diagonal element and its draw given a standard
    Gamma(shape, 1) variate g (Gibbs.scala:41-49,72-77; InverseGamma.scala:14)."""
    ss = _a(ss).ravel()
    k = ss.size
    g = _a(g).ravel()
    shape, rate, draw = np.empty(k), np.empty(k), np.empty(k)
    cnt = _a(count).ravel() if count is not None else None
    lib().oracle_gibbs_invgamma(k, C.c_double(prior_shape), C.c_double(prior_scale),
                                _p(cnt) if cnt is not None else None, C.c_double(count_all),
                                _p(ss), _p(g), _p(shape), _p(rate), _p(draw))
    return dict(shape=shape, rate=rate, draw=draw)


def inverse_wishart(n, psi, scatter, A):
    """InverseWishart(nu + T, psi + scatter).draw for an injected Bartlett factor A
    (GibbsWishart.scala:16-35, InverseWishart.scala:17-25).  Column-major flats."""
    W, sc = np.empty(n * n), np.empty(n * n)
    st = lib().oracle_inverse_wishart(n, _p(_a(psi)), _p(_a(scatter)) if scatter is not None else None,
                                      _p(_a(A)), _p(sc), _p(W))
    return dict(W=W, scale=sc, status=st)


def ar_filter(phi, mu, sigma, times, v, y, ou=False):
    """FilterAr / FilterOu.filterUnivariate (FilterAr.scala:15-47, FilterOu.scala:7-45)."""
    times, v, y = _a(times), _a(v), _a(y)
    T = times.size
    out = {k: np.empty(T + 1) for k in ("time", "m", "C", "a", "R")}
    lib().oracle_ar_filter(int(ou), T, C.c_double(phi), C.c_double(mu), C.c_double(sigma),
                           _p(times), _p(v), _p(y),
                           *(_p(out[k]) for k in ("time", "m", "C", "a", "R")))
    return out


def ar_backward_sample(phi, filt, z, ou=False):
    """FilterAr / FilterOu.univariateSample (FilterAr.scala:56-75, FilterOu.scala:47-71)."""
    T = filt["m"].size - 1
    theta = np.empty(T + 1)
    lib().oracle_ar_backward_sample(int(ou), T, C.c_double(phi), _p(filt["time"]), _p(filt["m"]),
                                    _p(filt["C"]), _p(filt["a"]), _p(filt["R"]), _p(_a(z)),
                                    _p(theta))
    return theta


def conjugate_filter(n, F, G, W, m0, C0, prior_shape, prior_scale, times, y):
    """ConjugateFilter(prior, advanceState).filter for p = 1 (ConjugateFilter.scala:23-94)."""
    times = _a(times)
    T = times.size
    F, f_tv, G, g_tv = _model(F, G, n, 1, T)
    out = dict(m=np.empty((T + 1, n)), C=np.empty((T + 1, n * n)), shape=np.empty(T + 1),
               scale=np.empty(T + 1))
    st = lib().oracle_conjugate_filter(n, T, _p(F), f_tv, _p(G), g_tv, _p(_a(W)), _p(_a(m0)),
                                       _p(_a(C0)), C.c_double(prior_shape), C.c_double(prior_scale),
                                       _p(times), _p(_a(y)), _p(out["m"]), _p(out["C"]),
                                       _p(out["shape"]), _p(out["scale"]))
    out["status"] = st
    return out
