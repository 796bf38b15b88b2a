"""C-ABI contract on a GPU box: error codes mirror the reference's exceptions, host-buffer calls
go through the slab pipeline for every entry point, nothing launches on API misuse."""
import ctypes as C

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


def _problem(capi, **kw):
    F = np.array([1.0]); G = np.array([1.0]); one = np.array([1.0]); y = np.zeros(4)
    base = dict(B=1, T=4, n=1, p=1, layout=capi.SERIES_MAJOR, mem=capi.HOST, keep_init=1, F=F, G=G,
                times=None, V=one, W=one, m0=np.zeros(1), C0=one, y=y)
    base.update(kw)
    keep = [v for v in base.values() if isinstance(v, np.ndarray)]
    return capi.make_problem(**base), keep


def test_error_codes(eng):
    from bayesian_dlms_b200 import _capi as capi
    lib, h = capi.load(), eng.ctx.handle
    out = np.zeros(8)
    ko = capi.KfOut(); ko.m = out.ctypes.data
    n0 = eng.ctx.launch_count()
    pr, keep = _problem(capi, T=0)
    assert lib.bdlm_kf_filter(h, pr, ko, None) == capi.E_EMPTY      # NoSuchElementException
    pr, keep = _problem(capi, n=49)   # BDLM_MAX_N = 48
    assert lib.bdlm_kf_filter(h, pr, ko, None) == capi.E_ARG
    pr, keep = _problem(capi, layout=7)
    assert lib.bdlm_kf_filter(h, pr, ko, None) == capi.E_ARG
    pr, keep = _problem(capi, y=None)
    assert lib.bdlm_kf_filter(h, pr, ko, None) == capi.E_ARG
    pr, keep = _problem(capi, keep_init=0)
    z = np.zeros(8); th = np.zeros(8)
    assert lib.bdlm_ffbs(h, pr, z.ctypes.data, th.ctypes.data, None, None, None) == capi.E_ARG
    assert b"keep_init" in lib.bdlm_last_error(h)
    assert lib.bdlm_kf_filter(None, pr, ko, None) == capi.E_ARG
    assert eng.ctx.launch_count() == n0, "API misuse must not launch anything"


def test_host_buffers_for_every_entry_point(eng):
    """mem = HOST for FFBS / SVD / log-likelihood / statistics, small staging cap so that the
    batch is really cut into several slabs."""
    import oracle
    from bayesian_dlms_b200 import Model, SERIES_MAJOR, dlm
    oracle.build()
    eng.ctx.set_staging_bytes(1 << 20)
    try:
        mod, V, W, m0, C0 = H.seasonal13()
        B, T, n, p = 150, 12, 13, 1
        rng = np.random.default_rng(4)
        times = np.arange(1, T + 1.0)
        y = np.stack([H.simulate(mod, V, W, m0, C0, times, rng, 0.1) for _ in range(B)])
        z = rng.standard_normal((B, T + 1, n))
        model = Model.build(mod, T=T)
        params = dict(V=V, W=W, m0=m0, C0=C0)
        F, _, G, _, _, _ = dlm.materialise(mod, times)
        out = eng.ffbs(model, params, y, z, layout=SERIES_MAJOR, stats=True)
        svd = eng.ffbs(model, params, y, z, layout=SERIES_MAJOR, svd=True)
        ll = eng.loglik(model, params, y, layout=SERIES_MAJOR)
        gs = eng.gibbs_stats(model, y, out["theta"], layout=SERIES_MAJOR)
        assert isinstance(out["theta"], np.ndarray) and (out["status"] == 0).all()
        for b in (0, 77, B - 1):
            o = oracle.ffbs(n, p, F, G, dlm.cm(V), dlm.cm(W), m0, dlm.cm(C0), times, y[b], z[b])
            assert np.array_equal(out["theta"][b], o["theta"])
            s = oracle.svd_ffbs(n, p, F, G, dlm.cm(V), dlm.cm(W), m0, dlm.cm(C0), times, y[b], z[b])
            assert np.array_equal(svd["theta"][b], s["theta"])
            l = oracle.loglik(n, p, F, G, dlm.cm(V), dlm.cm(W), m0, dlm.cm(C0), times, y[b])
            assert abs(ll["innovations"][b] - l["innovations"]) <= 1e-9 * abs(l["innovations"])
            assert np.array_equal(gs["ssw"][b], out["ssw"][b])
            assert np.array_equal(gs["scatter"][b], out["scatter"][b])
    finally:
        eng.ctx.set_staging_bytes(8 << 30)


def test_time_major_host_buffers_use_strided_copies(eng):
    import oracle
    from bayesian_dlms_b200 import Model, TIME_MAJOR, dlm
    oracle.build()
    eng.ctx.set_staging_bytes(1 << 20)
    try:
        mod, V, W, m0, C0 = H.second_order()
        B, T = 3000, 20
        rng = np.random.default_rng(5)
        y = rng.standard_normal((T, 1, B)).cumsum(axis=0)
        out = eng.filter_smooth(Model.build(mod, T=T), dict(V=V, W=W, m0=m0, C0=C0),
                                np.ascontiguousarray(y), layout=TIME_MAJOR)
        F, _, G, _, n, p = dlm.materialise(mod, np.arange(1, T + 1.0))
        for b in (0, 1500, B - 1):
            o = oracle.kf_filter(n, p, F, G, dlm.cm(V), dlm.cm(W), m0, dlm.cm(C0), np.arange(1, T + 1.0),
                                 y[:, :, b])
            s = oracle.rts_smooth(n, G, o)
            assert np.array_equal(out["m"][:, :, b], o["m"]) and np.array_equal(out["S"][:, :, b], s["S"])
    finally:
        eng.ctx.set_staging_bytes(8 << 30)


def test_distinct_contexts_run_concurrently_from_two_threads():
    """include/bdlm.h threading contract: a context is single-threaded-at-a-time, distinct contexts
    are fully concurrent (the reference's `mapAsync(nChains)` callers, Streaming.scala:162-173).
    Two host threads, one Engine each, hammer the same GPU; every result equals the
    single-threaded one."""
    import threading

    import torch
    from bayesian_dlms_b200 import Engine, Model, SERIES_MAJOR, dlm
    rng = np.random.default_rng(0)
    B, T = 257, 120
    mod = dlm.polynomial(2)
    params = dict(V=[[3.0]], W=np.diag([2.0, 1.0]), m0=np.zeros(2), C0=100.0 * np.eye(2))
    y = rng.standard_normal((B, T, 1)).cumsum(axis=1)       # host buffers: the library stages them
    z = rng.standard_normal((B, T + 1, 2))
    model = Model.build(mod, T=T)
    ref_eng = Engine(0)
    ref_fs = ref_eng.filter_smooth(model, params, y, layout=SERIES_MAJOR)
    ref_th = ref_eng.ffbs(model, params, y, z, layout=SERIES_MAJOR)["theta"]
    errors = []

    def worker(seed):
        try:
            eng = Engine(0)
            for _ in range(6):
                fs = eng.filter_smooth(model, params, y, layout=SERIES_MAJOR)
                th = eng.ffbs(model, params, y, z, layout=SERIES_MAJOR)["theta"]
                assert np.array_equal(fs["s"], ref_fs["s"]) and np.array_equal(fs["S"], ref_fs["S"])
                assert np.array_equal(th, ref_th)
        except Exception as ex:  # surfaced in the main thread
            errors.append(repr(ex))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    assert not errors, errors


def test_device_call_then_host_call_share_the_arena_safely(eng):
    """A device-mode call returns without synchronising and keeps using the context's workspace
    (spill, transposes); a host-mode call right behind it stages its slabs through the same arena
    on other streams.  The second must not start before the first has drained."""
    import oracle
    import torch
    from bayesian_dlms_b200 import Model, SERIES_MAJOR
    oracle.build()
    mod, V, W, m0, C0 = H.second_order()
    rng = np.random.default_rng(3)
    B, T = 40_000, 200
    y = rng.standard_normal((B, T, 1)).cumsum(axis=1)
    model = Model.build(mod, T=T)
    params = dict(V=V, W=W, m0=m0, C0=C0)
    yd = torch.from_numpy(y).cuda()
    for _ in range(3):
        d = eng.filter_smooth(model, params, yd, layout=SERIES_MAJOR, want=("s", "S"))   # async
        h = eng.filter_smooth(model, params, y[:64].copy(), layout=SERIES_MAJOR, want=("s", "S"))
        eng.sync()
        F, _, G, _, n, p = __import__("bayesian_dlms_b200").dlm.materialise(mod, np.arange(1, T + 1.0))
        cm = oracle.oracle.cm
        for b in (0, 63, B - 1):
            o = oracle.kf_filter(n, p, F, G, cm(V), cm(W), m0, cm(C0), np.arange(1, T + 1.0), y[b])
            s = oracle.rts_smooth(n, G, o)
            assert np.array_equal(d["S"][b].cpu().numpy(), s["S"]), b
            if b < 64:
                assert np.array_equal(h["S"][b], s["S"]), b


def test_filter_last_with_rank_deficient_W_reports_no_failure(eng):
    """W = diag(s2, 0) (trend / seasonal DLMs): the last filtered state and the innovations
    log-likelihood are well defined; only the transition density N(m_t; G m_{t-1}, W dt) is not,
    and it is evaluated -- and its status bits raised -- only when asked for."""
    import torch
    from bayesian_dlms_b200 import Model, SERIES_MAJOR
    from bayesian_dlms_b200 import _capi as capi
    mod, V, W, m0, C0 = H.second_order()
    W = np.diag([2.0, 0.0])
    rng = np.random.default_rng(4)
    for n_, model_mod, Wm in ((2, mod, W), (13, H.seasonal13()[0], np.diag([0.0] + [0.1] * 12))):
        Vm, m0m, C0m = np.array([[1.0]]), np.zeros(n_), np.eye(n_)
        T, B = 30, 5
        y = torch.from_numpy(rng.standard_normal((B, T, 1))).cuda()
        model = Model.build(model_mod, T=T)
        out = eng.filter_last(model, dict(V=Vm, W=Wm, m0=m0m, C0=C0m), y, layout=SERIES_MAJOR)
        eng.sync()
        assert int(out["status"].max()) == 0, n_
        full = eng.filter_last(model, dict(V=Vm, W=Wm, m0=m0m, C0=C0m), y, layout=SERIES_MAJOR,
                               loglik=True)
        eng.sync()
        assert int(full["status"].max()) & (capi.ST_SINGULAR | capi.ST_NOTPD), n_
        assert np.array_equal(out["m"].cpu().numpy(), full["m"].cpu().numpy())


def test_series_major_transposes_beyond_65535_row_tiles(eng):
    """Series-major register-kernel calls transpose through workspace; B > 65535 * 32 series used
    to exceed grid.y."""
    import torch
    from bayesian_dlms_b200 import Model, SERIES_MAJOR, TIME_MAJOR
    mod, V, W, m0, C0 = H.second_order()
    B, T = 65535 * 32 + 4097, 3
    g = torch.Generator(device="cuda").manual_seed(1)
    y = torch.randn((B, T, 1), generator=g, device="cuda", dtype=torch.float64)
    model = Model.build(mod, T=T)
    params = dict(V=V, W=W, m0=m0, C0=C0)
    a = eng.filter_smooth(model, params, y, layout=SERIES_MAJOR, want=("m", "S"))
    b = eng.filter_smooth(model, params, y.permute(1, 2, 0).contiguous(), layout=TIME_MAJOR,
                          want=("m", "S"))
    eng.sync()
    assert torch.equal(a["S"], b["S"].permute(2, 0, 1)) and torch.equal(a["m"], b["m"].permute(2, 0, 1))
