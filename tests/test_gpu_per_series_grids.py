"""Per-series time grids, covariates and ragged batches through the C ABI (BDLM_PS_TIMES / _F / _G).

``Data(time, observation)`` is per series in the reference (Dlm.scala:94): `Dlm.regression` has
F_t = (1, x_t) per series (Dlm.scala:159-169) and the AqMesh example per-sensor irregular times
(AqMeshExample.scala:86-127).  One batched call must equal the reference recursion (oracle) run
series by series on each series' own grid -- bit for bit."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import torch
    assert torch.cuda.is_available()
    from bayesian_dlms_b200 import default_engine
    return default_engine(0)


def _exact(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    ok = (a == b) | (np.isnan(a) & np.isnan(b))
    assert ok.all(), f"{what}: {np.sum(~ok)} of {a.size} differ, max rel {H.rel_err(a[~ok], b[~ok])}"


def _np(x):
    return x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)


def _sm(x, layout):
    """-> series-major (B, rows, k) numpy"""
    from bayesian_dlms_b200 import TIME_MAJOR
    x = _np(x)
    return x.transpose(2, 0, 1) if layout == TIME_MAJOR else x


def _case(name, rng, B, T):
    from bayesian_dlms_b200 import dlm
    times = np.cumsum(rng.choice([0.5, 1.0, 2.0, 3.25], (B, T)), axis=1) + rng.uniform(0, 5, (B, 1))
    if name == "trend":        # n = 2, p = 1: register kernels; G does not depend on dt
        mod, V, W, m0, C0 = H.second_order()
        mods = [mod] * B
    elif name == "regression":  # per-series covariates: F_t = (1, x_t), G = I
        xs = rng.standard_normal((B, T, 1))
        mods = [dlm.Dlm(lambda t, xb=xs[b], tb=times[b]: np.r_[1.0, xb[np.searchsorted(tb, t)]].reshape(2, 1),
                        lambda dt: np.eye(2)) for b in range(B)]
        V, W, m0, C0 = np.array([[1.5]]), np.diag([0.3, 0.1]), np.zeros(2), 10 * np.eye(2)
    elif name == "seasonal":    # n = 13: G(dt) differs by series on irregular grids
        mod, V, W, m0, C0 = H.seasonal13()
        mods = [mod] * B
    else:                        # correlated n = p = 8, G = I
        mod, V, W, m0, C0 = H.correlated8()
        mods = [mod] * B
    y = np.stack([H.simulate(mods[b], V, W, m0, C0, times[b], rng, missing=0.1) for b in range(B)])
    return mods, V, W, m0, C0, times, y


@pytest.mark.parametrize("name", ["trend", "regression", "seasonal", "correlated"])
@pytest.mark.parametrize("layout_name", ["series", "time"])
@pytest.mark.parametrize("where", ["device", "host"])
def test_batch_on_per_series_grids_equals_series_by_series_oracle(eng, name, layout_name, where):
    import oracle
    import torch
    from bayesian_dlms_b200 import (SERIES_MAJOR, TIME_MAJOR, build_batch_model, dlm, model_to_device)
    layout = SERIES_MAJOR if layout_name == "series" else TIME_MAJOR
    rng = np.random.default_rng(sum(map(ord, name + layout_name)))
    B, T = (7, 25) if name in ("seasonal", "correlated") else (67, 40)
    mods, V, W, m0, C0, times, y = _case(name, rng, B, T)
    n, p = len(m0), V.shape[0]
    model = build_batch_model(mods, times, layout=layout)
    assert "times" in model.per_series
    assert ("F" in model.per_series) == (name == "regression")
    assert ("G" in model.per_series) == (name == "seasonal")
    z = rng.standard_normal((B, T + 1, n))
    lay = (lambda a: np.ascontiguousarray(a.transpose(1, 2, 0))) if layout == TIME_MAJOR else np.ascontiguousarray
    yy, zz = lay(y), lay(z)
    if where == "device":
        model = model_to_device(model, "cuda")
        yy, zz = torch.from_numpy(yy).cuda(), torch.from_numpy(zz).cuda()
    params = dict(V=V, W=W, m0=m0, C0=C0)
    fs = eng.filter_smooth(model, params, yy, layout=layout)
    fb = eng.ffbs(model, params, yy, zz, layout=layout, stats=True)
    ll = eng.loglik(model, params, yy, layout=layout)
    sv = eng.ffbs(model, params, yy, zz, layout=layout, svd=True) if name != "regression" else None
    eng.sync()
    assert int(_np(fs["status"]).max()) == 0 and int(_np(fb["status"]).max()) == 0
    cm = oracle.oracle.cm
    for b in range(B):
        F, _, G, _, _, _ = dlm.materialise(mods[b], times[b])
        o = oracle.kf_filter(n, p, F, G, cm(V), cm(W), m0, cm(C0), times[b], y[b])
        for k in ("m", "C", "a", "R", "f", "Q"):
            _exact(_sm(fs[k], layout)[b][1:], o[k][1:], "%s series %d" % (k, b))
        s = oracle.rts_smooth(n, G, o)
        _exact(_sm(fs["s"], layout)[b], s["s"], "s")
        _exact(_sm(fs["S"], layout)[b], s["S"], "S")
        th = oracle.ffbs(n, p, F, G, cm(V), cm(W), m0, cm(C0), times[b], y[b], z[b])
        _exact(_sm(fb["theta"], layout)[b], th["theta"], "theta")
        st = oracle.gibbs_stats(n, p, F, G, th["time"], y[b], th["theta"])
        _exact(_np(fb["ssw"])[:, b] if layout == TIME_MAJOR else _np(fb["ssw"])[b], st["ssw"], "ssw")
        l = oracle.loglik(n, p, F, G, cm(V), cm(W), m0, cm(C0), times[b], y[b])
        assert abs(float(_np(ll["innovations"])[b]) - l["innovations"]) <= 1e-9 * abs(l["innovations"])
        assert abs(float(_np(ll["transition"])[b]) - l["transition"]) <= 1e-9 * abs(l["transition"])
        if sv is not None:
            osv = oracle.svd_ffbs(n, p, F, G, cm(V), cm(W), m0, cm(C0), times[b], y[b], z[b])
            _exact(_sm(sv["theta"], layout)[b], osv["theta"], "svd theta")


def test_ragged_batch_by_padding(eng):
    """Series of different LENGTHS in one call: a short series is padded at the end with its last
    time and NaN observations -- dt = 0 passes the state through (KalmanFilter.scala:277-279) and
    an all-missing update leaves it unchanged (:67-69), so its first len + 1 rows are exactly the
    unpadded run and the smoother's backward recursion starts from the same terminal state."""
    import oracle
    import torch
    from bayesian_dlms_b200 import SERIES_MAJOR, build_batch_model, dlm, model_to_device
    mod, V, W, m0, C0 = H.second_order()
    rng = np.random.default_rng(12)
    B, T = 33, 50
    lens = rng.integers(5, T + 1, B)
    lens[0] = T
    times = np.cumsum(rng.choice([1.0, 2.0], (B, T)), axis=1)
    y = np.stack([H.simulate(mod, V, W, m0, C0, times[b], rng, missing=0.05) for b in range(B)])
    for b in range(B):
        times[b, lens[b]:] = times[b, lens[b] - 1]
        y[b, lens[b]:] = np.nan
    model = model_to_device(build_batch_model(mod, times), "cuda")
    out = eng.filter_smooth(model, dict(V=V, W=W, m0=m0, C0=C0), torch.from_numpy(y).cuda(),
                            layout=SERIES_MAJOR)
    eng.sync()
    cm = oracle.oracle.cm
    for b in range(B):
        L = int(lens[b])
        F, _, G, _, n, p = dlm.materialise(mod, times[b, :L])
        o = oracle.kf_filter(n, p, F, G, cm(V), cm(W), m0, cm(C0), times[b, :L], y[b, :L])
        s = oracle.rts_smooth(n, G, o)
        _exact(_np(out["m"])[b, :L + 1], o["m"], "m")
        _exact(_np(out["C"])[b, :L + 1], o["C"], "C")
        _exact(_np(out["s"])[b, L], s["s"][L], "terminal s")
        # the padded tail repeats the terminal state
        _exact(_np(out["m"])[b, L:], np.repeat(o["m"][L:L + 1], T - L + 1, 0), "padded m")
        assert H.rel_err(_np(out["s"])[b, :L + 1], s["s"]) < 1e-9


def test_batched_overloads_of_the_mirror_api(eng):
    """`yss.map(ys => KalmanFilter.filterDlm(mod, ys, p))` and filter + backwardsSmoother as ONE GPU
    call (INTEGRATION.md batched overloads): series with their own irregular grids, their own
    lengths and their own DlmParameters equal the one-series mirror calls exactly."""
    from bayesian_dlms_b200 import Data, DlmParameters, KalmanFilter, Smoothing, polynomial
    rng = np.random.default_rng(21)
    mod = polynomial(2)
    yss, ps = [], []
    for b in range(9):
        T = int(rng.integers(8, 30))
        tm = np.cumsum(rng.choice([0.5, 1.0, 2.5], T)) + rng.uniform(0, 3)
        obs = rng.standard_normal(T).cumsum()
        yss.append([Data(float(tm[t]), np.array([None if rng.random() < 0.1 else float(obs[t])], dtype=object))
                    for t in range(T)])
        ps.append(DlmParameters(v=[[float(rng.uniform(1, 4))]], w=np.diag(rng.uniform(0.5, 2, 2)),
                                m0=rng.standard_normal(2), c0=10 * np.eye(2)))
    batch = KalmanFilter.filterDlmBatch(mod, yss, ps)
    filt, sm = KalmanFilter.filterSmoothBatch(mod, yss, ps)
    for b in range(len(yss)):
        one = KalmanFilter.filterDlm(mod, yss[b], ps[b])
        assert len(batch[b]) == len(one) == len(yss[b])
        for x, y_ in zip(batch[b], one):
            assert x.time == y_.time and np.array_equal(x.mt, y_.mt) and np.array_equal(x.ct, y_.ct)
            assert np.array_equal(x.ft, y_.ft) and np.array_equal(x.qt, y_.qt)
        f1 = KalmanFilter.filter(mod, yss[b], ps[b])
        s1 = Smoothing.backwardsSmoother(mod)(f1)
        assert len(sm[b]) == len(s1)
        assert np.array_equal(sm[b][-1].mean, s1[-1].mean)     # terminal state through the padding
        for x, y_ in zip(sm[b], s1):
            assert H.rel_err(x.mean, y_.mean) < 1e-9 and H.rel_err(x.covariance, y_.covariance) < 1e-9
