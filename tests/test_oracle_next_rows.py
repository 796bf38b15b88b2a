"""Oracle pins for the "next" rows of SURVEY.md section 8(f) (CPU, no GPU needed).

* AR(1) scalar filter: the reference's committed ``ar_dlm.csv -> ar_dlm_filtered.csv``
  (``FilterArDlm``, examples/src/main/scala/dlm/ar.scala:47-60) -- bit exact.
* Conjugate filter: ``first_order_dlm_conjugate_filtered.csv`` (``ConjFilter``,
  FirstOrderDlm.scala:144-172) -- bit exact on (m, C, E[V], Var[V]).
* Conjugate draws: closed-form checks of the posterior arithmetic (Gibbs.scala:41-49,72-77;
  GibbsWishart.scala:16-35; InverseWishart.scala:17-25) against numpy/LAPACK.
"""
import numpy as np
import pytest

import helpers as H
import oracle


def _ar_golden():
    rows = H.read_csv("ar_dlm.csv")
    times = np.array([float(r[0]) for r in rows])
    y = np.array([float(r[1]) for r in rows])
    g = H.read_csv("ar_dlm_filtered.csv")
    return times, y, np.array([[float(v) for v in r] for r in g])


def test_ar_filter_reproduces_reference_csv_bit_exact():
    times, y, gold = _ar_golden()
    out = oracle.ar_filter(0.8, 1.0, 0.3, times, np.full(times.size, 0.5), y)
    assert gold.shape == (times.size + 1, 3)
    assert np.array_equal(out["time"], gold[:, 0])
    assert np.array_equal(out["m"], gold[:, 1])
    assert np.array_equal(out["C"], gold[:, 2])


def test_ar_filter_is_the_dlm_filter_with_a_mean_shift():
    """AR(1) around mu == the generic DLM filter on (y - mu) with G = phi, W = sigma^2."""
    rng = np.random.default_rng(3)
    T, phi, mu, sig = 200, 0.9, -0.7, 0.4
    y = rng.standard_normal(T) + mu
    y[rng.random(T) < 0.1] = np.nan
    v = 0.5 + rng.random(T)
    ar = oracle.ar_filter(phi, mu, sig, np.arange(1.0, T + 1), v, y)
    # generic filter, one step at a time is awkward with varying v: use a numpy scalar recursion
    m, C = mu, sig * sig / (1 - phi * phi)
    for t in range(T):
        a, R = mu + phi * (m - mu), phi * phi * C + sig * sig
        if np.isnan(y[t]):
            m, C = a, R
        else:
            k = R / (R + v[t])
            m, C = a + k * (y[t] - a), k * v[t]
        assert abs(ar["m"][t + 1] - m) <= 1e-14 * max(1, abs(m))
        assert abs(ar["C"][t + 1] - C) <= 1e-14 * max(1, abs(C))


def test_ou_filter_and_sampler_moments():
    """OU transition moments (FilterOu.scala:12-18) and the backward sampler's conditional
    moments (FilterOu.scala:47-60) against a direct numpy evaluation."""
    rng = np.random.default_rng(5)
    T, phi, mu, sig = 50, 0.3, 1.2, 0.8
    times = np.cumsum(rng.uniform(0.2, 2.0, T))
    y = rng.standard_normal(T)
    v = np.full(T, 0.7)
    f = oracle.ar_filter(phi, mu, sig, times, v, y, ou=True)
    assert f["time"][0] == times[0] and f["m"][0] == mu        # t0 = head time, dt_1 = 0
    assert f["C"][0] == sig * sig / phi * phi                  # sic (FilterOu.scala:36)
    m, C, tp = f["m"][0], f["C"][0], f["time"][0]
    for t in range(T):
        dt = times[t] - tp
        a = mu + np.exp(-phi * dt) * (m - mu)
        R = np.exp(-2 * phi * dt) * C + sig ** 2 * (1 - np.exp(-2 * phi * dt)) / (2 * phi)
        k = R / (R + v[t])
        m, C, tp = a + k * (y[t] - a), k * v[t], times[t]
        assert np.isclose(f["a"][t + 1], a, rtol=1e-13) and np.isclose(f["R"][t + 1], R, rtol=1e-13)
        assert np.isclose(f["m"][t + 1], m, rtol=1e-13) and np.isclose(f["C"][t + 1], C, rtol=1e-13)
    z = rng.standard_normal(T + 1)
    th = oracle.ar_backward_sample(phi, f, z, ou=True)
    assert th[T] == f["m"][T] + np.sqrt(f["C"][T]) * z[T]
    for t in range(T - 1, -1, -1):
        ph = np.exp(-phi * (f["time"][t + 1] - f["time"][t]))
        mean = f["m"][t] + f["C"][t] * ph / f["R"][t + 1] * (th[t + 1] - f["a"][t + 1])
        cov = f["C"][t] - f["C"][t] ** 2 * ph ** 2 / f["R"][t + 1]
        assert np.isclose(th[t], mean + np.sqrt(cov) * z[t], rtol=1e-12, atol=1e-14)


def test_ar_backward_sampler_with_zero_noise_is_the_conditional_mean():
    times, y, _ = _ar_golden()
    times, y = times[:300], y[:300]
    f = oracle.ar_filter(0.8, 1.0, 0.3, times, np.full(300, 0.5), y)
    th = oracle.ar_backward_sample(0.8, f, np.zeros(301))
    assert th[300] == f["m"][300]
    # with z = 0 the draw is the RTS-smoothed mean of the scalar model
    s = f["m"][300]
    for t in range(299, -1, -1):
        s = f["m"][t] + f["C"][t] * 0.8 / f["R"][t + 1] * (s - f["a"][t + 1])
        assert np.isclose(th[t], s, rtol=1e-13)


def test_conjugate_filter_reproduces_reference_csv_bit_exact():
    times, y, _ = H.first_order_golden()
    gold = np.array([[float(v) for v in r] for r in H.read_csv("first_order_dlm_conjugate_filtered.csv")])
    out = oracle.conjugate_filter(1, [1.0], [1.0], [3.0], [0.0], [100.0], 3.0, 4.0, times, y[:, 0])
    assert out["status"] == 0
    assert gold.shape == (times.size + 1, 5)
    mean = out["scale"] / (out["shape"] - 1)
    var = (out["scale"] * out["scale"]) / ((out["shape"] - 1) * (out["shape"] - 1) * (out["shape"] - 2))
    assert np.array_equal(out["m"][:, 0], gold[:, 1])
    assert np.array_equal(out["C"][:, 0], gold[:, 2])
    assert np.array_equal(mean, gold[:, 3])
    assert np.array_equal(var, gold[:, 4])


def test_invgamma_posterior_arithmetic():
    ss = np.array([3.5, 0.25, 11.0])
    cnt = np.array([10.0, 7.0, 0.0])
    g = np.array([2.0, 0.5, 4.0])
    o = oracle.gibbs_invgamma(5.0, 4.0, ss, g, count=cnt)
    assert np.array_equal(o["shape"], 5.0 + cnt * 0.5)
    assert np.array_equal(o["rate"], 4.0 + ss * 0.5)
    assert np.array_equal(o["draw"], 1.0 / ((1.0 / o["rate"]) * g))
    o = oracle.gibbs_invgamma(17.0, 4.0, ss, g, count_all=2000.0)
    assert np.array_equal(o["shape"], np.full(3, 17.0 + 1000.0))


def test_inverse_wishart_matches_lapack():
    rng = np.random.default_rng(11)
    n = 5
    psi, scatter = H.spd(rng, n), H.spd(rng, n, 3.0)
    A = np.tril(rng.standard_normal((n, n)), -1) + np.diag(np.sqrt(rng.chisquare(12 - np.arange(n))))
    o = oracle.inverse_wishart(n, oracle.oracle.cm(psi), oracle.oracle.cm(scatter), oracle.oracle.cm(A))
    assert o["status"] == 0
    sc = psi + scatter
    l = np.linalg.cholesky(np.linalg.inv(sc))
    il, ia = np.linalg.inv(l), np.linalg.inv(A)
    want = il.T @ ia.T @ ia @ il
    got = o["W"].reshape(n, n).T
    assert H.rel_err(got, want) < 1e-10
    assert np.allclose(got, got.T, rtol=1e-12)
    assert np.all(np.linalg.eigvalsh(got) > 0)
    assert np.array_equal(o["scale"].reshape(n, n).T, sc)


def test_resumed_filter_equals_one_shot_and_forecast_matches_numpy():
    """t_init: folding KalmanFilter.step from a saved state (NoModel.scala:153-155) gives the same
    bits as filtering everything at once; Dlm.forecast (Dlm.scala:322-338) = the filter without
    observations, first dt = 0."""
    from bayesian_dlms_b200 import dlm
    rng = np.random.default_rng(12)
    mod = dlm.polynomial(1) + dlm.seasonal(24, 2)
    n, p = 5, 1
    V, W, m0, C0 = np.array([[1.5]]), np.diag(rng.uniform(0.1, 1, n)), rng.standard_normal(n), np.eye(n)
    times = np.cumsum(rng.choice([1.0, 2.0, 0.5], 80))
    y = H.simulate(mod, V, W, m0, C0, times, rng, missing=0.1)
    cm = oracle.oracle.cm
    F, f_tv, G, g_tv, _, _ = dlm.materialise(mod, times)
    full = oracle.kf_filter(n, p, F, G, cm(V), cm(W), m0, cm(C0), times, y)
    k = 33
    F2, _, G2, _, _, _ = dlm.materialise(mod, times[k:], t_init=times[k - 1])
    rest = oracle.kf_filter(n, p, F2, G2, cm(V), cm(W), full["m"][k], full["C"][k], times[k:], y[k:],
                            keep_init=False, t_init=times[k - 1])
    for key in ("m", "C", "a", "R", "f", "Q"):
        assert np.array_equal(rest[key], full[key][k + 1:]), key
    # forecast from the last state
    Hn = 6
    ft = times[-1] + np.arange(Hn)
    Ff, _, Gf, _, _, _ = dlm.materialise(mod, ft, t_init=times[-1])
    fc = oracle.kf_filter(n, p, Ff, Gf, cm(V), cm(W), full["m"][-1], full["C"][-1], ft,
                          np.full((Hn, 1), np.nan), keep_init=False, t_init=times[-1])
    m, C = full["m"][-1], full["C"][-1].reshape(n, n).T
    Fm = mod.f(0.0)
    assert np.allclose(fc["f"][0], Fm.T @ m, rtol=1e-13) and np.allclose(fc["Q"][0], Fm.T @ C @ Fm + V, rtol=1e-13)
    for h in range(1, Hn):
        Gm = mod.g(1.0)
        m, C = Gm @ m, Gm @ C @ Gm.T + W
        assert np.allclose(fc["f"][h], Fm.T @ m, rtol=1e-12)
        assert np.allclose(fc["Q"][h], Fm.T @ C @ Fm + V, rtol=1e-12)


def _np_kalman(mod, Vt, Wt, m0, C0, times, y, t_init=None):
    """Independent numpy Kalman filter (Joseph form, missing-aware), V and W per step."""
    m, C = m0.copy(), C0.copy()
    tp = times.min() - 1.0 if t_init is None else t_init
    out = []
    for t, tm in enumerate(times):
        dt = tm - tp
        F, G = mod.f(tm), mod.g(dt)
        a, R = (G @ m, G @ C @ G.T + Wt[t] * dt) if dt != 0 else (m, C)
        f, Q = F.T @ a, F.T @ R @ F + Vt[t]
        obs = ~np.isnan(y[t])
        if obs.any():
            Fo, Vo = F[:, obs], Vt[t][np.ix_(obs, obs)]
            Qo = Fo.T @ R @ Fo + Vo
            K = np.linalg.solve(Qo.T, Fo.T @ R.T).T
            m = a + K @ (y[t][obs] - Fo.T @ a)
            D = np.eye(len(m)) - K @ Fo.T
            C = D @ R @ D.T + K @ Vo @ K.T
        else:
            m, C = a, R
        out.append((m.copy(), C.copy(), a.copy(), R.copy(), f.copy(), Q.copy()))
        tp = tm
    return out


def test_time_varying_V_and_W_filter_matches_numpy():
    """oracle_kf_filter_tv (StudentTGibbs.filter, DlmFsvSystem.ffbs) against an independent numpy
    recursion: multivariate, irregular grid, partially missing observations."""
    from bayesian_dlms_b200 import dlm
    rng = np.random.default_rng(21)
    mod = dlm.polynomial(1) * dlm.polynomial(2)
    n, p, T = 3, 2, 40
    times = np.cumsum(rng.choice([0.5, 1.0, 2.0], T))
    Vt = np.stack([H.spd(rng, p, 2.0) for _ in range(T)])
    Wt = np.stack([H.spd(rng, n, 0.3) for _ in range(T)])
    m0, C0 = rng.standard_normal(n), H.spd(rng, n, 3.0)
    y = H.simulate(mod, np.eye(p), Wt[0], m0, C0, times, rng, missing=0.2)
    cm = oracle.oracle.cm
    F, _, G, _, _, _ = dlm.materialise(mod, times)
    o = oracle.kf_filter(n, p, F, G, np.stack([cm(v) for v in Vt]), np.stack([cm(w) for w in Wt]),
                         m0, cm(C0), times, y, keep_init=False, v_tv=True, w_tv=True)
    ref = _np_kalman(mod, Vt, Wt, m0, C0, times, y)
    for t, (m, C, a, R, f, Q) in enumerate(ref):
        assert np.allclose(o["m"][t], m, rtol=1e-10, atol=1e-12)
        assert np.allclose(o["C"][t].reshape(n, n).T, C, rtol=1e-9, atol=1e-12)
        assert np.allclose(o["R"][t].reshape(n, n).T, R, rtol=1e-9, atol=1e-12)
        assert np.allclose(o["Q"][t].reshape(p, p).T, Q, rtol=1e-9, atol=1e-12)


def test_conjugate_filter_matches_numpy_for_n2():
    """ConjugateFilter.step (ConjugateFilter.scala:53-94) for a two-dimensional state against a
    direct numpy evaluation, including the reference's `m = mt + k e` (previous mean, not at)."""
    from bayesian_dlms_b200 import dlm
    rng = np.random.default_rng(8)
    mod = dlm.polynomial(2)
    n, T = 2, 60
    W, m0, C0 = np.diag([0.5, 0.1]), np.array([0.3, -0.2]), 5.0 * np.eye(2)
    times = np.arange(1.0, T + 1)
    y = H.simulate(mod, np.array([[2.0]]), W, m0, C0, times, rng)[:, 0]
    cm = oracle.oracle.cm
    F, _, G, _, _, _ = dlm.materialise(mod, times)
    o = oracle.conjugate_filter(n, F, G, cm(W), m0, cm(C0), 4.0, 5.0, times, y)
    Fm, Gm = mod.f(1.0), mod.g(1.0)
    m, C, shape, scale = m0.copy(), C0.copy(), 4.0, 5.0
    for t in range(T):
        a, R = Gm @ m, Gm @ C @ Gm.T + W
        v = scale / (shape - 1)
        ft, qt = (Fm.T @ a)[0], (Fm.T @ R @ Fm)[0, 0] + v
        e = y[t] - ft
        K = (R @ Fm)[:, 0] / qt
        D = np.eye(n) - np.outer(K, Fm[:, 0])
        C = D @ R @ D.T + v * np.outer(K, K)
        m = m + K * e
        shape, scale = shape + 1, scale + v * e * e / qt
        assert np.allclose(o["m"][t + 1], m, rtol=1e-11) and np.allclose(o["C"][t + 1].reshape(n, n).T, C, rtol=1e-10)
        assert np.isclose(o["shape"][t + 1], shape) and np.isclose(o["scale"][t + 1], scale, rtol=1e-11)


def test_ar_ffbs_is_the_dlm_ffbs_of_the_equivalent_model():
    """An AR(1) state around mu = 0 with unit observation loading is the DLM F = 1, G = phi,
    W = sigma^2 with m0 = 0, C0 = sigma^2 / (1 - phi^2): FilterAr.ffbs (scalar formulas) and
    Smoothing.ffbs (matrix formulas + eigen draw) must agree to rounding on the same normals."""
    from bayesian_dlms_b200 import dlm
    rng = np.random.default_rng(31)
    T, phi, sig, v = 300, 0.85, 0.6, 0.9
    times = np.arange(1.0, T + 1)
    y = rng.standard_normal(T)
    y[rng.random(T) < 0.1] = np.nan
    z = rng.standard_normal(T + 1)
    f = oracle.ar_filter(phi, 0.0, sig, times, np.full(T, v), y)
    th = oracle.ar_backward_sample(phi, f, z)
    c0 = sig * sig / (1 - phi * phi)
    o = oracle.ffbs(1, 1, [1.0], [phi], [v], [sig * sig], [0.0], [c0], times, y.reshape(T, 1),
                    z.reshape(T + 1, 1))
    assert np.allclose(o["m"][:, 0], f["m"], rtol=1e-12, atol=1e-14)
    assert np.allclose(o["C"][:, 0], f["C"], rtol=1e-12, atol=1e-14)
    assert np.allclose(o["theta"][:, 0], th, rtol=1e-9, atol=1e-11)


# ------------------------------------------------------------------ f2 on the SVD path

@pytest.mark.parametrize("shape", [(2, 1), (3, 2)])
@pytest.mark.parametrize("tv", ["v", "w", "vw"])
def test_svd_filter_time_varying_params_vs_lapack_and_kalman(shape, tv):
    """DlmFsv.ffbsSvd (DlmFsv.scala:208-229) / DlmFsvSystem.ffbsSvd (DlmFsvSystem.scala:177-207):
    the SVD filter stepped with transformParams(p.copy(v = V_t)) [w = W_t] per observation, advance
    closure built from the transformed parameters.  Pins: (1) constant arrays reproduce the
    time-invariant oracle bit for bit; (2) the numpy + LAPACK restatement at 1e-8; (3) with the
    self-consistent closure and no partially missing rows the SVD filter is the Kalman filter in
    factored form, so the Kalman oracle with V_t / W_t must agree at 1e-7."""
    import oracle
    from oracle import lapack_flavour as lf
    from bayesian_dlms_b200 import dlm
    n, p = shape
    rng = np.random.default_rng(31 * n + p + len(tv))
    T = 40
    mod = dlm.polynomial(2) if p == 1 else dlm.polynomial(1) * dlm.polynomial(2)
    times = np.cumsum(rng.choice([1.0, 2.0], T))
    m0, C0 = rng.standard_normal(n), H.spd(rng, n, 4.0)
    v_tv, w_tv = "v" in tv, "w" in tv
    Vs = np.stack([H.spd(rng, p, 2.0) for _ in range(T if v_tv else 1)])
    Ws = np.stack([H.spd(rng, n, 0.4) for _ in range(T if w_tv else 1)])
    y = H.simulate(mod, Vs[0], Ws[0], m0, C0, times, rng)
    y[rng.random(T) < 0.15] = np.nan          # whole rows missing (Q6 needs no partial rows)
    F, _, G, _, _, _ = dlm.materialise(mod, times)
    cmT = lambda M: np.ascontiguousarray(M.transpose(0, 2, 1).reshape(M.shape[0], -1))  # noqa: E731
    z = rng.standard_normal((T + 1, n))
    o = oracle.svd_ffbs_tv(n, p, F, G, cmT(Vs), cmT(Ws), m0, dlm.cm(C0), times, y, z,
                           v_tv=v_tv, w_tv=w_tv, consistent=True)
    assert o["status"] == 0
    # (1) constant per-step arrays == the time-invariant entry point, bit for bit
    Vc, Wc = np.repeat(Vs[:1], T, 0), np.repeat(Ws[:1], T, 0)
    a = oracle.svd_ffbs_tv(n, p, F, G, cmT(Vc), cmT(Wc), m0, dlm.cm(C0), times, y, z,
                           v_tv=True, w_tv=True, consistent=True)
    b = oracle.svd_ffbs(n, p, F, G, dlm.cm(Vs[0]), dlm.cm(Ws[0]), m0, dlm.cm(C0), times, y, z,
                        consistent=True)
    for k in ("theta", "m", "dc", "uc", "a", "dr", "ur"):
        assert np.array_equal(a[k], b[k]), k
    # (2) LAPACK flavour
    Fs = lambda t: mod.f(times[t])                                      # noqa: E731
    tp = np.concatenate([[times.min() - 1.0], times])
    Gs = lambda t: mod.g(tp[t + 1] - tp[t])                             # noqa: E731
    Vf = lambda t: lf.sqrt_svd(Vs[t if v_tv else 0], inv=True)          # noqa: E731
    Wa = lambda t: lf.sqrt_svd(Ws[t if w_tv else 0])                    # noqa: E731
    ref = lf.svd_filter(Fs, Gs, Vf, Wa, m0, C0, times, y)
    assert H.rel_err(o["m"], np.stack([k["m"] for k in ref])) < 1e-8
    assert H.rel_err(o["dc"], np.stack([k["dc"] for k in ref])) < 1e-8
    # (3) Kalman filter with the same V_t, W_t
    kf = oracle.kf_filter(n, p, F, G, cmT(Vs) if v_tv else dlm.cm(Vs[0]),
                          cmT(Ws) if w_tv else dlm.cm(Ws[0]), m0, dlm.cm(C0), times, y,
                          v_tv=v_tv, w_tv=w_tv)
    assert H.rel_err(o["m"], kf["m"]) < 1e-7
    for r in range(T + 1):
        uc = dlm.from_cm(o["uc"][r], n, n)
        assert H.rel_err(uc @ np.diag(o["dc"][r] ** 2) @ uc.T, dlm.from_cm(kf["C"][r], n, n)) < 1e-6
    # the sampler's mean recursion (z = 0) against the LAPACK flavour, per-step sqrtW
    o0 = oracle.svd_ffbs_tv(n, p, F, G, cmT(Vs), cmT(Ws), m0, dlm.cm(C0), times, y, 0 * z,
                            v_tv=v_tv, w_tv=w_tv, consistent=True)
    if not w_tv:
        mom0 = lf.svd_sampler_moments(Gs, lf.sqrt_svd(Ws[0]), ref, o0["theta"])
        assert H.rel_err(o0["theta"], np.stack([x[0] for x in mom0])) < 1e-7
