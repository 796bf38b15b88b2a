"""Second, independent restatement of the reference path on numpy + scipy LAPACK.

TEST INFRASTRUCTURE ONLY.  Where oracle/bdlm_oracle.c replaces dgesv/dsyev/dgesdd by
hand-written LU / Jacobi routines (so the CUDA kernels can be bit-compared), this
module calls the LAPACK drivers Breeze 0.13.2 itself dispatches to (via netlib-java):
``\\`` -> dgesv, ``eigSym`` -> dsyev('V','L'), ``svd`` -> dgesdd.  It bounds how far the C
oracle can be from "what Breeze would produce" (tests assert <= 1e-9 relative on every
quantity that does not depend on LAPACK's implementation-defined eigenvector signs).

Matrices here are ordinary numpy (row, col) arrays; citations as in bdlm_oracle.c.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import lapack


def _solve(A, B):
    _, _, x, info = lapack.dgesv(np.asfortranarray(A), np.asfortranarray(B))
    assert info == 0
    return x


def _eigsym(A):
    w, v, info = lapack.dsyev(np.asfortranarray(A), lower=1)
    assert info == 0
    return w, v


def _svd(M):
    u, s, vt, info = lapack.dgesdd(np.asfortranarray(M), full_matrices=0)
    assert info == 0
    return s, vt


def kf_filter(Fs, Gs, V, W, m0, C0, times, y, keep_init=True):
    """Fs(t) -> (n,p), Gs(t) -> (n,n) callables on the observation index."""
    T = len(times)
    m, C = np.array(m0, float), np.array(C0, float)
    tprev = np.min(times) - 1.0
    out = []
    if keep_init:
        out.append(dict(time=tprev, m=m, C=C, a=m, R=C, f=None, Q=None))
    for t in range(T):
        F, G = Fs(t), Gs(t)
        dt = times[t] - tprev
        if dt == 0:
            a, R = m, C
        else:
            a = G @ m
            R = G @ C @ G.T + W * dt
        f = F.T @ a
        Q = F.T @ R @ F + V
        o = [i for i in range(F.shape[1]) if not np.isnan(y[t][i])]
        if not o:
            m, C = a, R
        else:
            Fm, Vm = F[:, o], V[np.ix_(o, o)]
            fm = Fm.T @ a
            Qm = Fm.T @ R @ Fm + Vm
            e = y[t][o] - fm
            K = _solve(Qm.T, Fm.T @ R.T).T
            m = a + K @ e
            D = np.eye(len(m)) - K @ Fm.T
            C = D @ R @ D.T + K @ Vm @ K.T
        tprev = times[t]
        out.append(dict(time=tprev, m=m, C=C, a=a, R=R, f=f, Q=Q))
    return out


def rts_smooth(Gs, filt, keep_init=True, textbook=False):
    rows = len(filt)
    s, S = filt[-1]["m"], filt[-1]["C"]
    out = [None] * rows
    out[-1] = (s, S)
    for r in range(rows - 2, -1, -1):
        G = Gs(r + 1 - int(keep_init))
        k, k1 = filt[r], filt[r + 1]
        B = _solve(k1["R"].T, G @ k["C"].T).T
        s_new = k["m"] + B @ (s - k1["a"])
        S_new = k["C"] - B @ (k1["R"] - S) @ (B.T if textbook else B)
        s, S = s_new, S_new
        out[r] = (s, S)
    return out


def sampler_moments(Gs, W, filt, theta, keep_init=True):
    """Per row r < rows-1: (h, H) of Smoothing.step given the *supplied* theta[r+1]."""
    rows = len(filt)
    out = [None] * rows
    out[-1] = (filt[-1]["m"], filt[-1]["C"])
    for r in range(rows - 2, -1, -1):
        G = Gs(r + 1 - int(keep_init))
        k, k1 = filt[r], filt[r + 1]
        dt = k1["time"] - k["time"]
        B = _solve(k1["R"].T, G @ k["C"].T).T
        h = k["m"] + B @ (theta[r + 1] - k1["a"])
        D = np.eye(len(h)) - B @ G
        H = D @ k["C"] @ D.T + B @ W * dt @ B.T
        H = (H + H.T) / 2.0
        out[r] = (h, H)
    return out


def eig_draw(mu, cov, z):
    w, v = _eigsym(cov)
    return mu + (v @ np.diag(np.sqrt(w))) @ z


def sqrt_svd(M, inv=False):
    s, vt = _svd(M)
    return np.diag(1.0 / np.sqrt(s) if inv else np.sqrt(s)) @ vt


def svd_filter(Fs, Gs, Vfac, Wadv, m0, C0, times, y, keep_init=True):
    """Vfac / Wadv: matrices, or callables t -> matrix (per-step V_t / W_t, DlmFsv.scala:208-229)."""
    T = len(times)
    Vfac_c, Wadv_c = Vfac, Wadv
    Vf = Vfac_c if callable(Vfac_c) else (lambda t: Vfac_c)
    Wa = Wadv_c if callable(Wadv_c) else (lambda t: Wadv_c)
    s, vt = _svd(np.array(C0, float))
    m, dc, uc = np.array(m0, float), np.sqrt(s), vt.T
    tprev = np.min(times) - 1.0
    out = []
    if keep_init:
        out.append(dict(time=tprev, m=m, dc=dc, uc=uc, a=m, dr=dc, ur=uc, f=Fs(0).T @ m))
    for t in range(T):
        F, G = Fs(t), Gs(t)
        dt = times[t] - tprev
        if dt == 0:
            a, dr, ur = m, dc, uc
        else:
            a = G @ m
            sv, vt = _svd(np.vstack([np.diag(dc) @ uc.T @ G.T, Wa(t) * np.sqrt(dt)]))
            ur, dr = vt.T, sv
        f = F.T @ a
        o = [i for i in range(F.shape[1]) if not np.isnan(y[t][i])]
        if not o:
            m, dc, uc = a, dr, ur
        else:
            Vm, Fm = Vf(t)[np.ix_(o, o)], F[:, o]
            sv, vt = _svd(np.vstack([Vm @ Fm.T @ ur, np.diag(1.0 / dr)]))
            uc = ur @ vt.T
            e = y[t][o] - Fm.T @ a
            fv = Fm @ Vm.T @ Vm
            dc = 1.0 / sv
            X = np.diag(dc) @ uc.T
            m = a + (X.T @ X @ fv) @ e
        tprev = times[t]
        out.append(dict(time=tprev, m=m, dc=dc, uc=uc, a=a, dr=dr, ur=ur, f=f))
    return out


def svd_sampler_moments(Gs, sqrtW, filt, theta, keep_init=True):
    """Per row: (h, cov = uh diag(dh^2) uh^T) of SvdSampler.step given theta[r+1]."""
    rows = len(filt)
    last = filt[-1]
    out = [None] * rows
    out[-1] = (last["m"], last["uc"] @ np.diag(last["dc"] ** 2) @ last["uc"].T)
    for r in range(rows - 2, -1, -1):
        G = Gs(r + 1 - int(keep_init))
        k, k1 = filt[r], filt[r + 1]
        sv, vt = _svd(np.vstack([sqrtW @ G @ k["uc"], np.diag(1.0 / k["dc"])]))
        uh, dh = k["uc"] @ vt.T, 1.0 / sv
        gW = G.T @ sqrtW.T @ sqrtW
        du = np.diag(dh) @ uh.T
        h = k["m"] + du.T @ du @ gW @ (theta[r + 1] - k1["a"])
        out[r] = (h, du.T @ du)
    return out


def mvn_logpdf(x, mu, S):
    c = x - mu
    slv = _solve(S, c.reshape(-1, 1)).ravel()
    L = np.linalg.cholesky(S)
    return -(slv @ c) / 2.0 - (len(x) / 2.0 * np.log(2 * np.pi) + np.sum(np.log(np.diag(L))))


def transition_loglik(Gs, W, filt):
    ll = 0.0
    for r in range(1, len(filt)):
        dt = filt[r]["time"] - filt[r - 1]["time"]
        ll += mvn_logpdf(filt[r]["m"], Gs(r - 1) @ filt[r - 1]["m"], W * dt)
    return ll
