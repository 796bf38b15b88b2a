"""GPU parity of the "next" rows through the C ABI, against the oracle.

* scalar AR(1) filter / backward sampler (FilterAr.scala): BIT-EXACT, incl. the reference's own
  golden CSV (ar_dlm.csv -> ar_dlm_filtered.csv);
* OU filter / sampler (FilterOu.scala): relative 1e-9 (device exp() vs libm exp());
* conjugate filter (ConjugateFilter.scala): BIT-EXACT, incl. first_order_dlm_conjugate_filtered.csv;
* conjugate draws (Gibbs.scala:41-49,72-77; GibbsWishart.scala:16-35): BIT-EXACT given injected
  variates; Philox-generated draws checked statistically and for reproducibility;
* a device-resident Gibbs run recovers the simulation parameters.
"""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def eng():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from bayesian_dlms_b200 import default_engine
    return default_engine(0)


@pytest.fixture(scope="module")
def oracle():
    import oracle as o
    o.build()
    return o


def _exact(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    ok = (a == b) | (np.isnan(a) & np.isnan(b))
    assert ok.all(), f"{what}: {np.sum(~ok)} of {a.size} differ, max rel {H.rel_err(a[~ok], b[~ok])}"


def _cuda(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _ar_case(rng, B, T, missing=0.1):
    phi = rng.uniform(0.5, 0.98, B)
    mu = rng.normal(0, 1, B)
    sig = rng.uniform(0.1, 1.0, B)
    y = rng.standard_normal((T, B)) + mu
    y[rng.random((T, B)) < missing] = np.nan
    v = rng.uniform(0.3, 3.0, (T, B))
    return phi, mu, sig, y, v


# ------------------------------------------------------------------ AR(1)

def test_ar_filter_reproduces_reference_csv(eng):
    from bayesian_dlms_b200 import TIME_MAJOR
    rows = H.read_csv("ar_dlm.csv")
    y = np.array([float(r[1]) for r in rows])
    gold = np.array([[float(v) for v in r] for r in H.read_csv("ar_dlm_filtered.csv")])
    yb = np.repeat(y[:, None], 3, axis=1)           # three identical series
    out = eng.ar_filter(dict(phi=0.8, mu=1.0, sigma_eta=0.3), _cuda(yb), 0.5, layout=TIME_MAJOR)
    for b in range(3):
        _exact(out["m"][:, b].cpu().numpy(), gold[:, 1], "m")
        _exact(out["C"][:, b].cpu().numpy(), gold[:, 2], "C")


@pytest.mark.parametrize("layout_name", ["time", "series"])
@pytest.mark.parametrize("mem", ["device", "host"])
def test_ar_filter_and_ffbs_bit_exact(eng, oracle, layout_name, mem):
    from bayesian_dlms_b200 import SERIES_MAJOR, TIME_MAJOR
    rng = np.random.default_rng(21)
    B, T = 37, 203
    phi, mu, sig, y, v = _ar_case(rng, B, T)
    z = rng.standard_normal((T + 1, B))
    layout = TIME_MAJOR if layout_name == "time" else SERIES_MAJOR
    arr = (lambda x: x) if layout == TIME_MAJOR else (lambda x: np.ascontiguousarray(x.T))
    put = _cuda if mem == "device" else np.ascontiguousarray
    get = (lambda t: t.cpu().numpy()) if mem == "device" else (lambda t: np.asarray(t))
    sv = dict(phi=put(phi), mu=put(mu), sigma_eta=put(sig))
    f = eng.ar_filter(sv, put(arr(y)), put(arr(v)), layout=layout)
    s = eng.ar_ffbs(sv, put(arr(y)), put(arr(v)), put(arr(z)), layout=layout, want=("m", "C"))
    s2 = eng.ar_ffbs(sv, put(arr(y)), put(arr(v)), put(arr(z)), layout=layout)  # workspace spill
    eng.sync()
    times = np.arange(1.0, T + 1)
    for b in range(B):
        o = oracle.ar_filter(phi[b], mu[b], sig[b], times, v[:, b], y[:, b])
        th = oracle.ar_backward_sample(phi[b], o, z[:, b])
        col = (lambda a: get(a)[:, b]) if layout == TIME_MAJOR else (lambda a: get(a)[b])
        for k in ("m", "C", "a", "R"):
            _exact(col(f[k]), o[k], k)
        _exact(col(s["theta"]), th, "theta")
        _exact(col(s2["theta"]), th, "theta (spill in workspace)")
        _exact(col(s["m"]), o["m"], "ffbs m")


def test_ar_shared_params_and_per_step_v(eng, oracle):
    rng = np.random.default_rng(5)
    B, T = 9, 64
    y = rng.standard_normal((T, B)) + 1.0
    v = rng.uniform(0.5, 2.0, T)
    f = eng.ar_filter(dict(phi=0.9, mu=1.0, sigma_eta=0.2), _cuda(y), v)
    for b in range(B):
        o = oracle.ar_filter(0.9, 1.0, 0.2, np.arange(1.0, T + 1), v, y[:, b])
        _exact(f["m"][:, b].cpu().numpy(), o["m"], "m")
        _exact(f["R"][:, b].cpu().numpy(), o["R"], "R")


def test_ou_filter_and_ffbs(eng, oracle):
    rng = np.random.default_rng(8)
    B, T = 21, 150
    phi, mu, sig, y, v = _ar_case(rng, B, T)
    phi = rng.uniform(0.1, 1.0, B)
    times = np.cumsum(rng.uniform(0.1, 2.5, T))
    z = rng.standard_normal((T + 1, B))
    sv = dict(phi=_cuda(phi), mu=_cuda(mu), sigma_eta=_cuda(sig))
    f = eng.ar_filter(sv, _cuda(y), _cuda(v), times=times, ou=True)
    s = eng.ar_ffbs(sv, _cuda(y), _cuda(v), _cuda(z), times=times, ou=True)
    for b in range(B):
        o = oracle.ar_filter(phi[b], mu[b], sig[b], times, v[:, b], y[:, b], ou=True)
        th = oracle.ar_backward_sample(phi[b], o, z[:, b], ou=True)
        for k in ("m", "C", "a", "R"):
            assert H.rel_err(f[k][:, b].cpu().numpy(), o[k]) < TOL, k
        # Row 0: FilterOu.filterUnivariate starts AT the first observation time (FilterOu.scala:37),
        # so dt_1 = 0, R_1 = C_0 and the backward variance C_0 - C_0^2 / R_1 is 0 up to rounding:
        # the reference itself returns NaN there when it rounds negative.  Same on both sides.
        got = s["theta"][:, b].cpu().numpy()
        assert np.array_equal(np.isnan(got[1:]), np.isnan(th[1:])) and not np.isnan(th[1:]).any()
        assert H.rel_err(got[1:], th[1:]) < 1e-7, "theta"
        assert np.isnan(got[0]) == np.isnan(th[0]) or abs(got[0] - th[0]) < 1e-6


def test_ar_empty_series_is_an_error(eng):
    import torch
    from bayesian_dlms_b200._capi import BdlmError, E_EMPTY
    with pytest.raises(BdlmError) as ei:
        eng.ar_filter(dict(phi=0.8, mu=0.0, sigma_eta=1.0),
                      torch.zeros((0, 4), dtype=torch.float64, device="cuda"), 1.0)
    assert ei.value.code == E_EMPTY


# ------------------------------------------------------------------ conjugate filter

def test_conjugate_filter_reproduces_reference_csv(eng):
    from bayesian_dlms_b200 import Model, dlm
    times, y, _ = H.first_order_golden()
    gold = np.array([[float(v) for v in r] for r in H.read_csv("first_order_dlm_conjugate_filtered.csv")])
    model = Model.build(dlm.polynomial(1), times=times)
    yb = np.repeat(y[:, :, None], 2, axis=2)
    out = eng.conjugate_filter(model, dict(W=[[3.0]], m0=[0.0], C0=[[100.0]]), 3.0, 4.0, _cuda(yb))
    assert int(out["status"].max()) == 0
    shape, scale = out["shape"][:, 0, 1].cpu().numpy(), out["scale"][:, 0, 1].cpu().numpy()
    _exact(out["m"][:, 0, 1].cpu().numpy(), gold[:, 1], "m")
    _exact(out["C"][:, 0, 1].cpu().numpy(), gold[:, 2], "C")
    _exact(scale / (shape - 1), gold[:, 3], "E[V]")
    _exact((scale * scale) / ((shape - 1) * (shape - 1) * (shape - 2)), gold[:, 4], "Var[V]")


@pytest.mark.parametrize("n", [1, 2, 3])
def test_conjugate_filter_bit_exact(eng, oracle, n):
    from bayesian_dlms_b200 import Model, SERIES_MAJOR, dlm
    rng = np.random.default_rng(40 + n)
    B, T = 11, 120
    mod = dlm.polynomial(n)
    W = np.diag(rng.uniform(0.1, 1.0, n))
    m0, C0 = rng.standard_normal(n), 10.0 * np.eye(n)
    times = np.cumsum(rng.choice([1.0, 1.0, 2.0, 0.5], T))
    y = np.stack([H.simulate(mod, np.array([[2.0]]), W, m0, C0, times, rng)[:, 0] for _ in range(B)])
    model = Model.build(mod, times=times)
    out = eng.conjugate_filter(model, dict(W=W, m0=m0, C0=C0), 5.0, 6.0, _cuda(y[:, :, None]),
                               layout=SERIES_MAJOR, want=("m", "C", "a", "R", "f", "Q"))
    F, G = model.F, model.G
    assert not model.g_tv, "polynomial G does not depend on dt"
    for b in range(B):
        o = oracle.conjugate_filter(n, F, G, oracle.oracle.cm(W), m0, oracle.oracle.cm(C0), 5.0, 6.0, times, y[b])
        _exact(out["m"][b].cpu().numpy(), o["m"], "m")
        _exact(out["C"][b].cpu().numpy(), o["C"], "C")
        _exact(out["shape"][b, :, 0].cpu().numpy(), o["shape"], "shape")
        _exact(out["scale"][b, :, 0].cpu().numpy(), o["scale"], "scale")


# ------------------------------------------------------------------ conjugate draws

@pytest.mark.parametrize("dims", [(13, 5, 3), (3, 37, 7)])
def test_gibbs_draw_injected_bit_exact(eng, oracle, dims):
    from bayesian_dlms_b200 import SERIES_MAJOR
    rng = np.random.default_rng(77)
    B, n, p = dims
    T = 250
    stats = dict(ssy=rng.uniform(1, 50, (B, p)), ny=rng.integers(100, T, (B, p)).astype(float),
                 ssw=rng.uniform(1, 50, (B, n)),
                 scatter=np.stack([oracle.oracle.cm(H.spd(rng, n, 5.0)) for _ in range(B)]))
    gv, gw = rng.gamma(60.0, 1.0, (B, p)), rng.gamma(130.0, 1.0, (B, n))
    A = np.stack([oracle.oracle.cm(np.tril(rng.standard_normal((n, n)), -1) +
                                   np.diag(np.sqrt(rng.chisquare(T + 10 - np.arange(n))))) for _ in range(B)])
    psi = H.spd(rng, n)
    dstats = {k: _cuda(v) for k, v in stats.items()}
    # d-inverse-gamma
    out = eng.gibbs_draw(n, p, T, dstats, dict(v_shape=5.0, v_scale=4.0, w_shape=17.0, w_scale=4.0),
                         layout=SERIES_MAJOR, inject=dict(gamma_v=_cuda(gv), gamma_w=_cuda(gw)),
                         want_shape_rate=True)
    # inverse Wishart
    outw = eng.gibbs_draw(n, p, T, dstats, dict(v_shape=5.0, v_scale=4.0, w_nu=10.0, w_psi=psi),
                          layout=SERIES_MAJOR, inject=dict(gamma_v=_cuda(gv), bartlett=_cuda(A)))
    assert int(out["status"].max()) == 0 and int(outw["status"].max()) == 0
    for b in range(B):
        ov = oracle.gibbs_invgamma(5.0, 4.0, stats["ssy"][b], gv[b], count=stats["ny"][b])
        ow = oracle.gibbs_invgamma(17.0, 4.0, stats["ssw"][b], gw[b], count_all=float(T))
        _exact(out["V"][b].cpu().numpy(), oracle.oracle.cm(np.diag(ov["draw"])), "V")
        _exact(out["W"][b].cpu().numpy(), oracle.oracle.cm(np.diag(ow["draw"])), "W")
        _exact(out["v_shape_rate"][b].cpu().numpy(), np.concatenate([ov["shape"], ov["rate"]]), "V shape|rate")
        _exact(out["w_shape_rate"][b].cpu().numpy(), np.concatenate([ow["shape"], ow["rate"]]), "W shape|rate")
        iw = oracle.inverse_wishart(n, oracle.oracle.cm(psi), stats["scatter"][b], A[b])
        assert iw["status"] == 0
        _exact(outw["W"][b].cpu().numpy(), iw["W"], "inverse Wishart W")
        _exact(outw["V"][b].cpu().numpy(), oracle.oracle.cm(np.diag(ov["draw"])), "V (wishart call)")


def test_gibbs_draw_philox_moments_and_reproducibility(eng):
    import torch
    from bayesian_dlms_b200 import TIME_MAJOR
    B, n, p, T = 200_000, 2, 1, 100
    ones = lambda k, v: torch.full((k, B), float(v), dtype=torch.float64, device="cuda")  # noqa: E731
    psi = np.array([[2.0, 0.3], [0.3, 1.0]])
    scatter = np.array([[30.0, 5.0], [5.0, 20.0]])
    stats = dict(ssy=ones(p, 40.0), ny=ones(p, 80.0), ssw=ones(n, 12.0),
                 scatter=torch.from_numpy(np.ascontiguousarray(scatter.T).ravel()).cuda()[:, None].repeat(1, B).contiguous())
    prior = dict(v_shape=5.0, v_scale=4.0, w_shape=17.0, w_scale=4.0)
    a = eng.gibbs_draw(n, p, T, stats, prior, layout=TIME_MAJOR, seed=123, sweep=0)
    b = eng.gibbs_draw(n, p, T, stats, prior, layout=TIME_MAJOR, seed=123, sweep=0)
    c = eng.gibbs_draw(n, p, T, stats, prior, layout=TIME_MAJOR, seed=123, sweep=1)
    assert torch.equal(a["V"], b["V"]) and torch.equal(a["W"], b["W"])
    assert not torch.equal(a["V"], c["V"])
    # InverseGamma(shape, rate): mean rate / (shape - 1), variance mean^2 / (shape - 2)
    sh, rt = 5.0 + 40.0, 4.0 + 20.0
    v = a["V"][0].cpu().numpy()
    assert abs(v.mean() - rt / (sh - 1)) < 4 * (rt / (sh - 1)) / np.sqrt((sh - 2) * B) + 1e-12
    assert abs(v.var() / ((rt / (sh - 1)) ** 2 / (sh - 2)) - 1) < 0.05
    shw, rtw = 17.0 + 50.0, 4.0 + 6.0
    w = a["W"].cpu().numpy()
    assert np.all(w[1] == 0) and np.all(w[2] == 0)
    for k in (0, 3):
        assert abs(w[k].mean() / (rtw / (shw - 1)) - 1) < 0.005
    assert abs(np.corrcoef(w[0], w[3])[0, 1]) < 0.01 and abs(np.corrcoef(v, w[0])[0, 1]) < 0.01
    # inverse Wishart: E[W] = (psi + scatter) / (nu + T - d - 1)
    iw = eng.gibbs_draw(n, p, T, stats, dict(v_shape=5.0, v_scale=4.0, w_nu=6.0, w_psi=psi),
                        layout=TIME_MAJOR, seed=9, sweep=3)
    assert int(iw["status"].max()) == 0
    Wm = iw["W"].cpu().numpy().mean(axis=1).reshape(2, 2).T
    want = (psi + scatter) / (6.0 + T - 2 - 1)
    assert np.allclose(Wm, want, rtol=0.01), (Wm, want)
    Ws = iw["W"].cpu().numpy()
    assert np.array_equal(Ws[1], Ws[2]) or np.allclose(Ws[1], Ws[2], rtol=1e-12)   # symmetric


def test_device_resident_gibbs_recovers_parameters(eng):
    """GibbsSampling.sample (Gibbs.scala:153-180) batched: first-order DLM, V = 2, W = 3."""
    import torch
    from bayesian_dlms_b200 import Model, TIME_MAJOR, dlm, gibbs
    rng = np.random.default_rng(1)
    B, T = 64, 400
    mod = dlm.polynomial(1)
    y = np.stack([H.simulate(mod, np.array([[2.0]]), np.array([[3.0]]), np.zeros(1), np.eye(1),
                             np.arange(1.0, T + 1), rng, missing=0.05)[:, 0] for _ in range(B)], axis=1)
    model = Model.build(mod, T=T)
    res = gibbs.sample(eng, model, _cuda(y[:, None, :]),
                       dict(v_shape=3.0, v_scale=4.0, w_shape=3.0, w_scale=6.0),
                       dict(V=[[1.0]], W=[[1.0]], m0=[0.0], C0=[[10.0]]), 300, seed=4, layout=TIME_MAJOR)
    torch.cuda.synchronize()
    assert int(res["status"].max()) == 0
    v = res["V"][100:, 0, :].mean().item()
    w = res["W"][100:, 0, :].mean().item()
    assert abs(v - 2.0) < 0.35 and abs(w - 3.0) < 0.5, (v, w)


# ------------------------------------------------------------------ time-varying V_t (f2)

@pytest.mark.parametrize("n,p", [(2, 1), (3, 2), (13, 1)])
@pytest.mark.parametrize("shared", [True, False])
def test_time_varying_V_filter_and_ffbs(eng, oracle, n, p, shared):
    """StudentTGibbs.filter / sampleState (StudentTGibbs.scala:100-136): KalmanFilter.step with
    params.copy(v = V_t) at step t, then Smoothing.sampleDlm.  Bit-exact against the oracle."""
    from bayesian_dlms_b200 import Model, SERIES_MAJOR, dlm
    rng = np.random.default_rng(100 * n + p + int(shared))
    B, T = 5, 60
    if n == 13:
        mod, _, W, m0, C0 = H.seasonal13()
    else:
        mod = dlm.polynomial(n) if p == 1 else dlm.polynomial(1) * dlm.polynomial(2)
        W, m0, C0 = H.spd(rng, n, 0.3), rng.standard_normal(n), H.spd(rng, n, 4.0)
    times = np.arange(1.0, T + 1)
    Vb = np.stack([np.stack([H.spd(rng, p, 2.0) for _ in range(T)]) for _ in range(1 if shared else B)])
    y = np.stack([H.simulate(mod, np.eye(p), W, m0, C0, times, rng, missing=0.1) for _ in range(B)])
    z = rng.standard_normal((B, T + 1, n))
    model = Model.build(mod, T=T)
    Vflat = np.ascontiguousarray(Vb.transpose(0, 1, 3, 2).reshape(Vb.shape[0], T, p * p))  # col-major rows
    if shared:
        params = dict(V=Vflat[0], W=W, m0=m0, C0=C0, v_tv=True)
    else:
        params = dict(V=_cuda(Vflat), W=W, m0=m0, C0=C0, v_tv=True, per_series=("V",))
    f = eng.filter(model, params, _cuda(y), layout=SERIES_MAJOR)
    s = eng.ffbs(model, params, _cuda(y), _cuda(z), layout=SERIES_MAJOR)
    fs = eng.filter_smooth(model, params, _cuda(y), layout=SERIES_MAJOR)
    assert int(f["status"].max()) == 0 and int(s["status"].max()) == 0
    cm = oracle.oracle.cm
    for b in range(B):
        Vt = Vflat[0 if shared else b]
        o = oracle.kf_filter(n, p, model.F, model.G, Vt, cm(W), m0, cm(C0), times, y[b], v_tv=True)
        for k in ("m", "C", "a", "R", "f", "Q"):
            _exact(f[k][b].cpu().numpy()[1:], o[k][1:], k)
            _exact(fs[k][b].cpu().numpy()[1:], o[k][1:], "fused " + k)
        sm = oracle.rts_smooth(n, model.G, o)
        _exact(fs["s"][b].cpu().numpy(), sm["s"], "s")
        _exact(fs["S"][b].cpu().numpy(), sm["S"], "S")
        th = oracle.ffbs(n, p, model.F, model.G, Vt, cm(W), m0, cm(C0), times, y[b], z[b], v_tv=True)
        _exact(s["theta"][b].cpu().numpy(), th["theta"], "theta")


# ------------------------------------------------------------------ reference-API mirrors

def test_reference_api_mirrors_of_the_next_rows(eng):
    """The calls as the reference's own apps make them: FilterArDlm (ar.scala:47-60), ConjFilter
    (FirstOrderDlm.scala:144-172), GibbsSampling.sample (Gibbs.scala:153-180)."""
    from bayesian_dlms_b200 import (ConjugateFilter, Data, DlmParameters, FilterAr, GibbsSampling,
                                    InverseGamma, SvParameters, polynomial)
    rows = H.read_csv("ar_dlm.csv")[:500]
    data = [(float(r[0]), float(r[1])) for r in rows]
    gold = np.array([[float(v) for v in r] for r in H.read_csv("ar_dlm_filtered.csv")[:501]])
    p = SvParameters(0.8, 1.0, 0.3)
    filtered = FilterAr.filterUnivariate(data, [0.5] * len(data), p)
    assert len(filtered) == len(data) + 1
    _exact([s.time for s in filtered], gold[:, 0], "time")
    _exact([s.mt for s in filtered], gold[:, 1], "mt")
    _exact([s.ct for s in filtered], gold[:, 2], "ct")
    sampled = FilterAr.ffbs(p, data, [0.5] * len(data), z=np.zeros(len(data) + 1))
    assert sampled[-1].sample == filtered[-1].mt and len(sampled) == len(filtered)

    times, y, _ = H.first_order_golden()
    obs = [Data(t, [v]) for t, v in zip(times, y[:, 0])]
    goldc = np.array([[float(v) for v in r] for r in H.read_csv("first_order_dlm_conjugate_filtered.csv")])
    pc = DlmParameters(v=2.0, w=3.0, m0=0.0, c0=100.0)
    out = ConjugateFilter(InverseGamma(3.0, 4.0)).filter(polynomial(1), obs, pc)
    _exact([s.kfState.mt[0] for s in out], goldc[:, 1], "conj m")
    _exact([s.kfState.ct[0, 0] for s in out], goldc[:, 2], "conj C")
    _exact([s.variance[0].mean for s in out], goldc[:, 3], "E[V]")
    _exact([s.variance[0].variance for s in out], goldc[:, 4], "Var[V]")
    assert out[0].kfState.ft is None

    chains = GibbsSampling.sample(polynomial(1), InverseGamma(3.0, 4.0), InverseGamma(3.0, 6.0),
                                  DlmParameters(v=1.0, w=1.0, m0=0.0, c0=10.0), obs[:300], 50,
                                  chains=4, seed=2)
    assert len(chains) == 4 and len(chains[0]) == 50
    assert all(c[-1].v[0, 0] > 0 and c[-1].w[0, 0] > 0 for c in chains)
    assert chains[0][-1].v[0, 0] != chains[1][-1].v[0, 0]


# ------------------------------------------------------------------ resume / forecast (f4)

def test_filter_from_saved_state_and_forecast(eng, oracle):
    from bayesian_dlms_b200 import Data, DlmParameters, KalmanFilter, dlm
    rng = np.random.default_rng(3)
    mod = dlm.polynomial(1) + dlm.seasonal(24, 2)
    n = 5
    V, W, m0, C0 = np.array([[1.5]]), np.diag(rng.uniform(0.1, 1, n)), rng.standard_normal(n), np.eye(n)
    times = np.cumsum(rng.choice([1.0, 2.0, 0.5], 60))
    y = H.simulate(mod, V, W, m0, C0, times, rng, missing=0.1)
    data = [Data(t, [None if np.isnan(v) else v for v in row]) for t, row in zip(times, y)]
    p = DlmParameters(V, W, m0, C0)
    full = KalmanFilter.filter(mod, data, p)              # T + 1 states
    k = 25
    rest = KalmanFilter.filterFrom(mod, full[k], data[k:], p)
    assert len(rest) == len(data) - k
    for a, b in zip(rest, full[k + 1:]):
        assert a.time == b.time
        _exact(a.mt, b.mt, "mt"); _exact(a.ct, b.ct, "ct"); _exact(a.at, b.at, "at")
        _exact(a.rt, b.rt, "rt"); _exact(a.ft, b.ft, "ft"); _exact(a.qt, b.qt, "qt")
    fc = KalmanFilter.forecast(mod, full[-1].mt, full[-1].ct, full[-1].time, p, 5)
    cm = oracle.oracle.cm
    ft = full[-1].time + np.arange(5.0)
    Ff, _, Gf, _, _, _ = dlm.materialise(mod, ft, t_init=full[-1].time)
    o = oracle.kf_filter(n, 1, Ff, Gf, cm(V), cm(W), full[-1].mt, cm(full[-1].ct), ft,
                         np.full((5, 1), np.nan), keep_init=False, t_init=full[-1].time)
    for h, (t, f, q) in enumerate(fc):
        assert t == ft[h]
        _exact(f, o["f"][h], "forecast mean"); _exact(q.ravel(), o["Q"][h], "forecast variance")
    # first forecast step is oneStepPrediction on the state itself (dt = 0)
    assert np.allclose(fc[0][1], mod.f(0.0).T @ full[-1].mt)


@pytest.mark.parametrize("shape", [(3, 2), (2, 1)])
@pytest.mark.parametrize("shared", [True, False])
def test_time_varying_W_filter_and_ffbs(eng, oracle, shared, shape):
    """DlmFsvSystem.ffbs (DlmFsvSystem.scala:137-167): KalmanFilter.step with params.copy(w = W_t)
    forward, Smoothing.step(model, W_t) for the transition t -> t + 1 backward; with V_t as well."""
    from bayesian_dlms_b200 import Model, SERIES_MAJOR, dlm
    rng = np.random.default_rng(77 + int(shared))
    B, T = 4, 45
    n, p = shape
    mod = dlm.polynomial(1) * dlm.polynomial(2) if p == 2 else dlm.polynomial(2)
    m0, C0 = rng.standard_normal(n), H.spd(rng, n, 4.0)
    times = np.cumsum(rng.choice([1.0, 2.0], T))
    nb = 1 if shared else B
    Wb = np.stack([np.stack([H.spd(rng, n, 0.4) for _ in range(T)]) for _ in range(nb)])
    Vb = np.stack([np.stack([H.spd(rng, p, 2.0) for _ in range(T)]) for _ in range(nb)])
    y = np.stack([H.simulate(mod, np.eye(p), Wb[0, 0], m0, C0, times, rng, missing=0.1) for _ in range(B)])
    z = rng.standard_normal((B, T + 1, n))
    model = Model.build(mod, times=times)
    flat = lambda M, k: np.ascontiguousarray(M.transpose(0, 1, 3, 2).reshape(M.shape[0], T, k))  # noqa: E731
    Wf, Vf = flat(Wb, n * n), flat(Vb, p * p)
    if shared:
        params = dict(V=Vf[0], W=Wf[0], m0=m0, C0=C0, v_tv=True, w_tv=True)
    else:
        params = dict(V=_cuda(Vf), W=_cuda(Wf), m0=m0, C0=C0, v_tv=True, w_tv=True, per_series=("V", "W"))
    f = eng.filter(model, params, _cuda(y), layout=SERIES_MAJOR)
    s = eng.ffbs(model, params, _cuda(y), _cuda(z), layout=SERIES_MAJOR)
    assert int(f["status"].max()) == 0 and int(s["status"].max()) == 0
    cm = oracle.oracle.cm
    for b in range(B):
        Vt, Wt = Vf[0 if shared else b], Wf[0 if shared else b]
        o = oracle.kf_filter(n, p, model.F, model.G, Vt, Wt, m0, cm(C0), times, y[b], v_tv=True, w_tv=True)
        for k in ("m", "C", "a", "R", "f", "Q"):
            _exact(f[k][b].cpu().numpy()[1:], o[k][1:], k)
        th = oracle.ffbs(n, p, model.F, model.G, Vt, Wt, m0, cm(C0), times, y[b], z[b], v_tv=True, w_tv=True)
        _exact(s["theta"][b].cpu().numpy(), th["theta"], "theta")


def test_gibbs_draw_host_buffers_time_major(eng, oracle):
    """Same call with HOST (numpy) buffers and the time-major layout: the library stages them."""
    from bayesian_dlms_b200 import TIME_MAJOR
    rng = np.random.default_rng(5)
    B, n, p, T = 9, 4, 2, 120
    stats = dict(ssy=rng.uniform(1, 50, (p, B)), ny=rng.integers(50, T, (p, B)).astype(float),
                 ssw=rng.uniform(1, 50, (n, B)), scatter=np.zeros((n * n, B)))
    gv, gw = rng.gamma(40.0, 1.0, (p, B)), rng.gamma(70.0, 1.0, (n, B))
    out = eng.gibbs_draw(n, p, T, stats, dict(v_shape=5.0, v_scale=4.0, w_shape=17.0, w_scale=4.0),
                         layout=TIME_MAJOR, inject=dict(gamma_v=gv, gamma_w=gw))
    assert isinstance(out["V"], np.ndarray) and out["V"].shape == (p * p, B)
    for b in range(B):
        ov = oracle.gibbs_invgamma(5.0, 4.0, stats["ssy"][:, b], gv[:, b], count=stats["ny"][:, b])
        ow = oracle.gibbs_invgamma(17.0, 4.0, stats["ssw"][:, b], gw[:, b], count_all=float(T))
        _exact(out["V"][:, b], oracle.oracle.cm(np.diag(ov["draw"])), "V")
        _exact(out["W"][:, b], oracle.oracle.cm(np.diag(ow["draw"])), "W")


@pytest.mark.parametrize("svd", [False, True])
def test_device_resident_gibbs_wishart(eng, svd):
    """GibbsWishart.sample (GibbsWishart.scala:64-80) / the SVD variant: FFBS + inverse-Wishart W +
    inverse-gamma V, all sweeps on the device; draws stay symmetric positive definite."""
    import torch
    from bayesian_dlms_b200 import Model, SERIES_MAJOR, dlm, gibbs
    rng = np.random.default_rng(9)
    B, T = 12, 150
    mod = dlm.polynomial(1) * dlm.polynomial(1)          # CorrelatedModel.scala:16-27 in miniature
    V = np.diag([1.0, 4.0])
    W = np.array([[0.75, 0.5], [0.5, 1.25]])
    y = np.stack([H.simulate(mod, V, W, np.zeros(2), np.eye(2), np.arange(1.0, T + 1), rng, missing=0.05)
                  for _ in range(B)])
    model = Model.build(mod, T=T)
    res = gibbs.sample(eng, model, _cuda(y), dict(v_shape=6.0, v_scale=5.0, w_nu=10.0, w_psi=np.eye(2)),
                       dict(V=V, W=W, m0=np.zeros(2), C0=np.eye(2)), 60, seed=3, layout=SERIES_MAJOR,
                       svd=svd)
    torch.cuda.synchronize()
    assert int(res["status"].max()) == 0
    Wc = res["W"].cpu().numpy()                           # (iters, 4, B) column-major 2 x 2
    assert np.allclose(Wc[:, 1], Wc[:, 2], rtol=1e-9)
    det = Wc[:, 0] * Wc[:, 3] - Wc[:, 1] * Wc[:, 2]
    assert (Wc[:, 0] > 0).all() and (det > 0).all()
    m = Wc[20:].mean(axis=(0, 2))
    assert 0.2 < m[0] < 2.5 and 0.3 < m[3] < 3.5, m


# ------------------------------------------------------------------ on-device RNG mode (z = None)

def test_ffbs_rng_mode_is_keyed_by_seed_sweep_series_row_component(eng):
    """bdlm_set_rng: with z = NULL the kernels draw Philox normals that depend only on (seed, sweep,
    global series, row, component): same key -> same path whatever the layout, batch split or
    kernel (warp, group, four-series SVD); another sweep -> another path; N(0,1) marginals."""
    import torch
    from scipy import stats as sst
    from bayesian_dlms_b200 import Model, SERIES_MAJOR, TIME_MAJOR, dlm
    rng = np.random.default_rng(2)
    # (a) n = 1: theta_T = m_T + sqrt(C_T) z  ->  recover z and test it
    B, T = 40_000, 3
    mod = dlm.polynomial(1)
    params = dict(V=[[2.0]], W=[[3.0]], m0=[0.0], C0=[[10.0]])
    y = _cuda(rng.standard_normal((T, 1, B)))
    model = Model.build(mod, T=T)
    eng.ctx.set_rng(11, 0)
    a = eng.ffbs(model, params, y, None, want_kf=("m", "C"))
    b = eng.ffbs(model, params, y, None, want_kf=("m", "C"))
    eng.ctx.set_rng(11, 1)
    c = eng.ffbs(model, params, y, None)
    assert torch.equal(a["theta"], b["theta"]) and not torch.equal(a["theta"], c["theta"])
    z = ((a["theta"][-1, 0] - a["m"][-1, 0]) / a["C"][-1, 0].sqrt()).cpu().numpy()
    assert abs(z.mean()) < 0.03 and abs(z.var() - 1) < 0.03
    assert sst.kstest(z, "norm").pvalue > 1e-3
    assert abs(np.corrcoef(z[:-1], z[1:])[0, 1]) < 0.02          # neighbouring series independent
    # (b) layout, batch split and kernel variant do not change a series' path
    for mk, svd in ((H.seasonal7 if hasattr(H, "seasonal7") else None, False), (H.correlated8, True),
                    (H.seasonal13, False)):
        if mk is None:
            continue
        mod, V, W, m0, C0 = mk()
        n, p, Bs, Ts = len(m0), V.shape[0], 7, 12
        ys = np.stack([H.simulate(mod, V, W, m0, C0, np.arange(1.0, Ts + 1), rng, missing=0.1) for _ in range(Bs)])
        model = Model.build(mod, T=Ts)
        pr = dict(V=V, W=W, m0=m0, C0=C0)
        eng.ctx.set_rng(5, 3)
        s1 = eng.ffbs(model, pr, _cuda(ys), None, layout=SERIES_MAJOR, svd=svd)["theta"].cpu().numpy()
        t1 = eng.ffbs(model, pr, _cuda(np.ascontiguousarray(ys.transpose(1, 2, 0))), None, layout=TIME_MAJOR,
                      svd=svd)["theta"].cpu().numpy().transpose(2, 0, 1)
        _exact(s1, t1, "layout")
        eng.ctx.set_rng(5, 3, first_series=4)
        s2 = eng.ffbs(model, pr, _cuda(ys[4:]), None, layout=SERIES_MAJOR, svd=svd)["theta"].cpu().numpy()
        _exact(s2, s1[4:], "batch split with first_series")
        eng.ctx.set_rng(5, 3)


# ------------------------------------------------------------------ Rao-Blackwellised step (f4)

def test_rao_blackwell_kf_step_over_a_particle_cloud(eng, oracle):
    """RaoBlackwellFilter.kfStep (RaoBlackwellFilter.scala:43-57) for every parameter particle in
    ONE call: a batch of B particles, each with its own (V, W, m, C), one observation (T = 1)
    arriving dt after the particles' common time -> (mt1, ct1) and the conditional likelihood
    (KalmanFilter.conditionalLikelihood, KalmanFilter.scala:138-153) used as the new weight."""
    from bayesian_dlms_b200 import Model, SERIES_MAJOR, dlm
    rng = np.random.default_rng(4)
    B, n, p = 300, 3, 2
    mod = dlm.polynomial(1) * dlm.polynomial(2)
    cm = oracle.oracle.cm
    Vs = np.stack([cm(H.spd(rng, p, 2.0)) for _ in range(B)])
    Ws = np.stack([cm(H.spd(rng, n, 0.5)) for _ in range(B)])
    ms = rng.standard_normal((B, n))
    Cs = np.stack([cm(H.spd(rng, n, 3.0)) for _ in range(B)])
    y = np.tile(np.array([[[0.7, np.nan]]]), (B, 1, 1))        # second sensor missing
    y[::2, 0, 1] = -1.3
    t_prev, t_now = 41.0, 43.5
    model = Model.build(mod, times=[t_now], t_init=t_prev)
    params = dict(V=_cuda(Vs), W=_cuda(Ws), m0=_cuda(ms), C0=_cuda(Cs), per_series=("V", "W", "m0", "C0"))
    f = eng.filter(model, params, _cuda(y), layout=SERIES_MAJOR, keep_init=False, want=("m", "C", "f", "Q"))
    ll = eng.loglik(model, params, _cuda(y), layout=SERIES_MAJOR)
    assert int(f["status"].max()) == 0
    for b in range(0, B, 7):
        o = oracle.kf_filter(n, p, model.F, model.G, Vs[b], Ws[b], ms[b], Cs[b], [t_now], y[b],
                             keep_init=False, t_init=t_prev)
        _exact(f["m"][b, 0].cpu().numpy(), o["m"][0], "mt1")
        _exact(f["C"][b, 0].cpu().numpy(), o["C"][0], "ct1")
        # conditional likelihood of the observed components
        obs = ~np.isnan(y[b, 0])
        Q = o["Q"][0].reshape(p, p).T[np.ix_(obs, obs)]
        e = (y[b, 0] - o["f"][0])[obs]
        want = -0.5 * (e @ np.linalg.solve(Q, e)) - 0.5 * (obs.sum() * np.log(2 * np.pi) + np.linalg.slogdet(Q)[1])
        got = float(ll["innovations"][b])
        assert abs(got - want) <= 1e-9 * max(1.0, abs(want)), (b, got, want)


@pytest.mark.parametrize("name", ["second_order", "seasonal7"])
def test_filter_last_and_streaming_resume(eng, oracle, name):
    """bdlm_kf_filter_last: only the final state (no per-step stores) == the last row of the full
    filter, bit for bit; feeding it back with t_init continues the filter exactly (the streaming
    pattern of NoModel.scala:153-155)."""
    from bayesian_dlms_b200 import Model, SERIES_MAJOR, dlm
    rng = np.random.default_rng(6)
    if name == "second_order":
        mod, V, W, m0, C0 = H.second_order()
    else:
        mod = dlm.polynomial(1) + dlm.seasonal(24, 3)
        V, W, m0, C0 = np.array([[1.0]]), np.diag([0.01, 0.2, 0.4, 0.5, 0.2, 0.1, 0.4]), np.zeros(7), np.eye(7)
    n, B, T, k = len(m0), 19, 70, 31
    y = np.stack([H.simulate(mod, V, W, m0, C0, np.arange(1.0, T + 1), rng, missing=0.1) for _ in range(B)])
    params = dict(V=V, W=W, m0=m0, C0=C0)
    full = eng.filter(Model.build(mod, T=T), params, _cuda(y), layout=SERIES_MAJOR, want=("m", "C"))
    last = eng.filter_last(Model.build(mod, T=k), params, _cuda(y[:, :k]), layout=SERIES_MAJOR, loglik=True)
    assert int(last["status"].max()) == 0
    _exact(last["m"].cpu().numpy(), full["m"][:, k].cpu().numpy(), "m_last")
    _exact(last["C"].cpu().numpy(), full["C"][:, k].cpu().numpy(), "C_last")
    ll = eng.loglik(Model.build(mod, T=k), params, _cuda(y[:, :k]), layout=SERIES_MAJOR)
    assert np.array_equal(last["innovations"].cpu().numpy(), ll["innovations"].cpu().numpy())
    # resume from the saved state
    model2 = Model.build(mod, times=np.arange(k + 1.0, T + 1), t_init=float(k))
    p2 = dict(V=V, W=W, m0=last["m"], C0=last["C"], per_series=("m0", "C0"))
    rest = eng.filter(model2, p2, _cuda(y[:, k:]), layout=SERIES_MAJOR, keep_init=False, want=("m", "C"))
    _exact(rest["m"].cpu().numpy(), full["m"][:, k + 1:].cpu().numpy(), "resumed m")
    _exact(rest["C"].cpu().numpy(), full["C"][:, k + 1:].cpu().numpy(), "resumed C")


# ------------------------------------------------------------------ f2 on the SVD path

@pytest.mark.parametrize("shape", [(2, 1), (3, 2), (8, 8), (13, 1)])   # <= 8: four-series kernel
@pytest.mark.parametrize("tv", ["v", "w", "vw"])
@pytest.mark.parametrize("shared", [True, False])
def test_svd_path_time_varying_V_W(eng, oracle, shape, tv, shared):
    """DlmFsv.ffbsSvd (DlmFsv.scala:208-229: SvdFilter.step with transformParams(p.copy(v = V_t)),
    sampler on ps.head.w) and DlmFsvSystem.ffbsSvd (DlmFsvSystem.scala:177-207: W_t forward and in
    SvdSampler.step of the transition t -> t + 1), both closures self-consistent (BDLM_SVD_CONSISTENT_W);
    the raw-W (quirk Q2) closure with per-step parameters as well.  Bit-exact against the oracle."""
    from bayesian_dlms_b200 import Model, SERIES_MAJOR, dlm
    n, p = shape
    rng = np.random.default_rng(1000 * n + 10 * p + len(tv) + int(shared))
    B, T = 6, 30
    if n == 13:
        mod = H.seasonal13()[0]
    elif n == 8:
        mod = H.correlated8()[0]
    else:
        mod = dlm.polynomial(2) if p == 1 else dlm.polynomial(1) * dlm.polynomial(2)
    m0, C0 = rng.standard_normal(n), H.spd(rng, n, 4.0)
    times = np.cumsum(rng.choice([1.0, 2.0], T))
    v_tv, w_tv = "v" in tv, "w" in tv
    nb = 1 if shared else B
    Vb = np.stack([np.stack([H.spd(rng, p, 2.0) for _ in range(T if v_tv else 1)]) for _ in range(nb)])
    Wb = np.stack([np.stack([H.spd(rng, n, 0.4) for _ in range(T if w_tv else 1)]) for _ in range(nb)])
    y = np.stack([H.simulate(mod, np.eye(p), Wb[0, 0], m0, C0, times, rng, missing=0.1) for _ in range(B)])
    z = rng.standard_normal((B, T + 1, n))
    model = Model.build(mod, times=times)
    flat = lambda M, k: np.ascontiguousarray(M.transpose(0, 1, 3, 2).reshape(M.shape[0], -1, k))  # noqa: E731
    Vf, Wf = flat(Vb, p * p), flat(Wb, n * n)
    params = dict(m0=m0, C0=C0, v_tv=v_tv, w_tv=w_tv)
    per = []
    for name, arr, is_tv in (("V", Vf, v_tv), ("W", Wf, w_tv)):
        if shared:
            params[name] = arr[0] if is_tv else arr[0, 0]
        else:
            params[name] = _cuda(arr if is_tv else arr[:, 0])
            per.append(name)
    if per:
        params["per_series"] = tuple(per)
    for consistent in (True, False):
        f = eng.svd_filter(model, params, _cuda(y), layout=SERIES_MAJOR, consistent_w=consistent)
        s = eng.ffbs(model, params, _cuda(y), _cuda(z), layout=SERIES_MAJOR, svd=True,
                     consistent_w=consistent, want_kf=("m", "dc"))
        assert int(f["status"].max()) == 0 and int(s["status"].max()) == 0
        for b in range(B):
            Vt = Vf[0 if shared else b] if v_tv else Vf[0 if shared else b][0]
            Wt = Wf[0 if shared else b] if w_tv else Wf[0 if shared else b][0]
            o = oracle.svd_ffbs_tv(n, p, model.F, model.G, Vt, Wt, m0, oracle.oracle.cm(C0), times,
                                   y[b], z[b], v_tv=v_tv, w_tv=w_tv, consistent=consistent)
            for k in ("m", "dc", "uc", "a", "dr", "ur"):
                _exact(f[k][b].cpu().numpy(), o[k], "%s consistent=%s" % (k, consistent))
            _exact(s["theta"][b].cpu().numpy(), o["theta"], "theta consistent=%s" % consistent)
            _exact(s["svd_m"][b].cpu().numpy(), o["m"], "ffbs m")


@pytest.mark.parametrize("shape", [(37, 7), (40, 3), (48, 2)])
def test_svd_path_beyond_32_states(eng, oracle, shape):
    """The SVD entry points at the reference's largest example (n = 37, p = 7 with missing sensors
    and fractional irregular times, AqMeshExample.scala:86-127) and up to BDLM_MAX_N = 48."""
    from bayesian_dlms_b200 import Model, SERIES_MAJOR, dlm
    n, p = shape
    rng = np.random.default_rng(n + p)
    B, T = 3, 12
    # p sensors loading on n states: a dense 0/1-ish F and a dt-dependent G (damped levels with a
    # weak coupling to the next state), so every irregular step has its own G
    Fm = (rng.random((n, p)) < 0.3).astype(float) + np.eye(n, p)
    mod = dlm.Dlm(lambda t: Fm,
                  lambda dt: np.exp(-0.05 * dt) * np.eye(n) + 0.1 * dt * np.eye(n, k=1))
    times = np.cumsum(rng.uniform(0.25, 1.5, T))
    V, W = H.spd(rng, p, 1.0), H.spd(rng, n, 0.2)
    m0, C0 = rng.standard_normal(n), H.spd(rng, n, 2.0)
    y = np.stack([H.simulate(mod, V, W, m0, C0, times, rng, missing=0.2) for _ in range(B)])
    z = rng.standard_normal((B, T + 1, n))
    model = Model.build(mod, times=times)
    assert model.n == n and model.p == p
    params = dict(V=V, W=W, m0=m0, C0=C0)
    f = eng.svd_filter(model, params, _cuda(y), layout=SERIES_MAJOR)
    s = eng.ffbs(model, params, _cuda(y), _cuda(z), layout=SERIES_MAJOR, svd=True)
    assert int(f["status"].max()) == 0 and int(s["status"].max()) == 0
    cm = oracle.oracle.cm
    for b in range(B):
        o = oracle.svd_ffbs(n, p, model.F, model.G, cm(V), cm(W), m0, cm(C0), times, y[b], z[b])
        for key in ("m", "dc", "uc", "a", "dr", "ur"):
            _exact(f[key][b].cpu().numpy(), o[key], key)
        _exact(s["theta"][b].cpu().numpy(), o["theta"], "theta")
