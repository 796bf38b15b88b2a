"""Drop-in mirror of the reference's operator API for the Kalman hot path.

Same names, argument meaning, result types and error behaviour as the Scala objects
``KalmanFilter``, ``Smoothing``, ``SvdFilter``, ``SvdSampler`` and the Gibbs sufficient
statistics, so the parity tests read like the reference's own tests.  Every call runs
on the GPU through libbdlm.so (batch of one series, host buffers); there is no CPU
implementation behind this module.  The Scala facade that binds the same C ABI from
the JVM is described in INTEGRATION.md.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _capi as capi
from . import dlm as _dlm
from .batch import Model, SERIES_MAJOR, default_engine
from .dlm import Data, Dlm, DlmParameters


class MatrixSingularException(ArithmeticError):
    """Breeze's exception for a zero pivot in ``\\`` (status bit BDLM_ST_SINGULAR)."""


class NotConvergedException(ArithmeticError):
    """Breeze's exception for eigSym/svd non-convergence (BDLM_ST_NOTCONVERGED)."""


def _raise_status(st: int):
    if st & capi.ST_SINGULAR:
        raise MatrixSingularException("singular matrix in solve")
    if st & capi.ST_NOTCONVERGED:
        raise NotConvergedException("Jacobi iteration did not converge")


@dataclass
class KfState:
    """``KfState`` (KalmanFilter.scala:22-30); ft/qt are None at the initial state."""
    time: float
    mt: np.ndarray
    ct: np.ndarray
    at: np.ndarray
    rt: np.ndarray
    ft: Optional[np.ndarray]
    qt: Optional[np.ndarray]


@dataclass
class SmoothingState:
    """``Smoothing.SmoothingState`` (Smoothing.scala:18-22)."""
    time: float
    mean: np.ndarray
    covariance: np.ndarray
    at1: np.ndarray
    rt1: np.ndarray


@dataclass
class SamplingState:
    """``SamplingState`` (Smoothing.scala:10-15)."""
    time: float
    sample: np.ndarray
    mean: np.ndarray
    cov: np.ndarray
    at1: np.ndarray
    rt1: np.ndarray


@dataclass
class SvdState:
    """``SvdState`` (SvdFilter.scala:7-14)."""
    time: float
    mt: np.ndarray
    dc: np.ndarray
    uc: np.ndarray
    at: np.ndarray
    dr: np.ndarray
    ur: np.ndarray
    ft: np.ndarray


def _prep(mod: Dlm, ys: Sequence[Data], p: DlmParameters):
    times, y = _dlm.flatten_data(ys)  # raises on empty data like t0.get
    model = Model.build(mod, times)
    params = dict(V=p.v, W=p.w, m0=p.m0, C0=p.c0)
    return model, params, times, np.ascontiguousarray(y.reshape(1, model.T, model.p))


def _row_times(times: np.ndarray, keep_init: bool):
    return np.concatenate([[times.min() - 1.0], times]) if keep_init else times


def _kf_states(model, out, tm, keep_init):
    n, p = model.n, model.p
    res = []
    for r in range(len(tm)):
        init = keep_init and r == 0
        res.append(KfState(
            float(tm[r]), out["m"][0, r].copy(), _dlm.from_cm(out["C"][0, r], n, n).copy(),
            out["a"][0, r].copy(), _dlm.from_cm(out["R"][0, r], n, n).copy(),
            None if init else out["f"][0, r].copy(),
            None if init else _dlm.from_cm(out["Q"][0, r], p, p).copy()))
    return res


def _pack_filtered(filtered: Sequence[KfState]):
    m = np.stack([s.mt for s in filtered])[None]
    C = np.stack([_dlm.cm(s.ct) for s in filtered])[None]
    a = np.stack([s.at for s in filtered])[None]
    R = np.stack([_dlm.cm(s.rt) for s in filtered])[None]
    return dict(m=np.ascontiguousarray(m), C=np.ascontiguousarray(C),
                a=np.ascontiguousarray(a), R=np.ascontiguousarray(R))


def _model_for_states(mod: Dlm, state_times: np.ndarray):
    """Model for a vector of filtered states whose first element may be the initial state:
    G between consecutive states is g(time[r+1] - time[r])."""
    # states[1:] play the role of the observations; state 0 supplies the previous time.
    # The smoother only needs G between consecutive states (a, R come with the KfStates).
    st = np.asarray(state_times, dtype=np.float64)
    F, f_tv, G, g_tv, n, p = _dlm.materialise_dts(mod, st[1:], np.diff(st))
    return Model(np.ascontiguousarray(F.ravel()), np.ascontiguousarray(G.ravel()), bool(f_tv),
                 bool(g_tv), n, p, int(st.size - 1), None)


def _prep_batch(mod, yss: Sequence[Sequence[Data]], ps):
    """Batched overloads (last rows): ``yss`` holds one ``Vector[Data]`` per
    series -- each with its OWN times (Dlm.scala:94) and, for filter / smoother calls, its own
    length -- ``mod`` one ``Dlm`` or one per series (regression covariates), ``ps`` one
    ``DlmParameters`` or one per series.  Short series are padded with ``None`` observations at
    their last time (dt = 0 passes the state through, KalmanFilter.scala:277-279)."""
    from .batch import build_batch_model
    B = len(yss)
    flat = [_dlm.flatten_data(ys) for ys in yss]
    lens = np.array([t.size for t, _ in flat])
    T, p = int(lens.max()), flat[0][1].shape[1]
    times = np.empty((B, T))
    y = np.full((B, T, p), np.nan)
    for b, (t, yy) in enumerate(flat):
        times[b, :t.size] = t
        times[b, t.size:] = t[-1]
        y[b, :t.size] = yy
    same_grid = not isinstance(mod, (list, tuple)) and all(np.array_equal(times[0], times[b]) for b in range(1, B))
    model = Model.build(mod, times[0]) if same_grid else build_batch_model(mod, times)
    if isinstance(ps, DlmParameters):
        params = dict(V=ps.v, W=ps.w, m0=ps.m0, C0=ps.c0)
    else:
        assert len(ps) == B
        params = dict(V=np.stack([_dlm.cm(q.v) for q in ps]), W=np.stack([_dlm.cm(q.w) for q in ps]),
                      m0=np.stack([q.m0 for q in ps]), C0=np.stack([_dlm.cm(q.c0) for q in ps]),
                      per_series=("V", "W", "m0", "C0"))
    return model, params, times, y, lens


class KalmanFilter:
    """``KalmanFilter`` companion-object entry points (KalmanFilter.scala)."""

    @staticmethod
    def filterBatch(mod, yss: Sequence[Sequence[Data]], ps, keep_init: bool = True) -> List[List[KfState]]:
        """``yss.map(ys => KalmanFilter(adv).filter(mod, ys, p))`` as ONE call on the GPU: every
        series on its own (irregular) time grid and of its own length, per-series parameters when
        ``ps`` is a sequence (the reference's "one model per sensor" loop, UoModel.scala:69-104)."""
        model, params, times, y, lens = _prep_batch(mod, yss, ps)
        out = default_engine().filter(model, params, y, layout=SERIES_MAJOR, keep_init=keep_init)
        res = []
        for b in range(len(yss)):
            _raise_status(int(out["status"][b]))
            L = int(lens[b]) + int(keep_init)
            one = {k: out[k][b:b + 1, :L] for k in ("m", "C", "a", "R", "f", "Q")}
            res.append(_kf_states(model, one, _row_times(times[b, :lens[b]], keep_init), keep_init))
        return res

    @staticmethod
    def filterDlmBatch(mod, yss, ps) -> List[List[KfState]]:
        """Batched ``KalmanFilter.filterDlm`` (:291-294): initial states dropped."""
        return KalmanFilter.filterBatch(mod, yss, ps, keep_init=False)

    @staticmethod
    def filterSmoothBatch(mod, yss, ps):
        """``filter`` followed by ``Smoothing.backwardsSmoother`` for every series in one fused call.
        Returns (filtered, smoothed) lists of lists."""
        model, params, times, y, lens = _prep_batch(mod, yss, ps)
        out = default_engine().filter_smooth(model, params, y, layout=SERIES_MAJOR)
        n = model.n
        filt, sm = [], []
        for b in range(len(yss)):
            _raise_status(int(out["status"][b]))
            L = int(lens[b]) + 1
            tm = _row_times(times[b, :lens[b]], True)
            one = {k: out[k][b:b + 1, :L] for k in ("m", "C", "a", "R", "f", "Q")}
            ks = _kf_states(model, one, tm, True)
            filt.append(ks)
            sm.append([SmoothingState(float(tm[r]), out["s"][b, r].copy(),
                                      _dlm.from_cm(out["S"][b, r], n, n).copy(), ks[r].at, ks[r].rt)
                       for r in range(L)])
        return filt, sm

    @staticmethod
    def filterDlm(mod: Dlm, ys: Sequence[Data], p: DlmParameters) -> List[KfState]:
        """``KalmanFilter.filterDlm`` (:291-294): T states, initial state dropped."""
        return KalmanFilter._run(mod, ys, p, keep_init=False)

    @staticmethod
    def filter(mod: Dlm, ys: Sequence[Data], p: DlmParameters) -> List[KfState]:
        """``KalmanFilter(advanceState(p, mod.g)).filter`` (Filter.scala:41-45): T+1 states."""
        return KalmanFilter._run(mod, ys, p, keep_init=True)

    @staticmethod
    def _run(mod, ys, p, keep_init):
        model, params, times, y = _prep(mod, ys, p)
        out = default_engine().filter(model, params, y, layout=SERIES_MAJOR, keep_init=keep_init)
        _raise_status(int(out["status"][0]))
        return _kf_states(model, out, _row_times(times, keep_init), keep_init)

    @staticmethod
    def filterFrom(mod: Dlm, state: KfState, ys: Sequence[Data], p: DlmParameters) -> List[KfState]:
        """``ys.foldLeft(state)(KalmanFilter(adv).step(mod, p))`` collected (KalmanFilter.scala:99-107,
        the "carry on from the last filtered state" pattern of NoModel.scala:153-155): the states
        after each new observation, starting from a saved ``KfState``."""
        times, y = _dlm.flatten_data(ys)
        model = Model.build(mod, times, t_init=state.time)
        params = dict(V=p.v, W=p.w, m0=state.mt, C0=state.ct)
        yb = np.ascontiguousarray(y.reshape(1, model.T, model.p))
        out = default_engine().filter(model, params, yb, layout=SERIES_MAJOR, keep_init=False)
        _raise_status(int(out["status"][0]))
        return _kf_states(model, out, times, False)

    @staticmethod
    def forecast(mod: Dlm, mt: np.ndarray, ct: np.ndarray, time: float, p: DlmParameters,
                 horizon: int):
        """``Dlm.forecast(mod, mt, ct, time, p).take(horizon)`` (Dlm.scala:322-338): the forecast
        mean and variance of the observation at ``time``, ``time + 1``, ...  A filter run without
        observations from the state (mt, ct) at ``time``: the first step has dt = 0
        (``oneStepPrediction`` on the state itself), the following ones are ``stepForecast``."""
        if horizon <= 0:
            return []
        times = float(time) + np.arange(horizon, dtype=np.float64)
        model = Model.build(mod, times, t_init=float(time))
        params = dict(V=p.v, W=p.w, m0=np.asarray(mt, float), C0=np.asarray(ct, float))
        yb = np.full((1, horizon, model.p), np.nan)
        out = default_engine().filter(model, params, yb, layout=SERIES_MAJOR, keep_init=False,
                                      want=("f", "Q"))
        _raise_status(int(out["status"][0]))
        pp = model.p
        return [(float(times[r]), out["f"][0, r].copy(), _dlm.from_cm(out["Q"][0, r], pp, pp).copy())
                for r in range(horizon)]

    @staticmethod
    def likelihood(mod: Dlm, ys: Sequence[Data]):
        """``KalmanFilter.likelihood(mod, ys)(p)`` (:299-306), curried like the reference."""
        def at(p: DlmParameters) -> float:
            model, params, times, y = _prep(mod, ys, p)
            out = default_engine().loglik(model, params, y, layout=SERIES_MAJOR)
            return float(out["transition"][0])
        return at

    @staticmethod
    def innovationsLikelihood(mod: Dlm, ys: Sequence[Data]):
        """Sum over t of ``KalmanFilter.conditionalLikelihood`` (:138-153)."""
        def at(p: DlmParameters) -> float:
            model, params, times, y = _prep(mod, ys, p)
            out = default_engine().loglik(model, params, y, layout=SERIES_MAJOR)
            return float(out["innovations"][0])
        return at


class Smoothing:
    """``Smoothing`` object entry points (Smoothing.scala)."""

    @staticmethod
    def backwardsSmoother(mod: Dlm, w: Optional[np.ndarray] = None):
        """``Smoothing.backwardsSmoother(mod)(kfState)`` (:57-64).  The smoother recursion
        never reads W; it is only needed when a_{t+1}, R_{t+1} must be recomputed, which
        does not happen here because the KfStates carry them."""
        def run(filtered: Sequence[KfState]) -> List[SmoothingState]:
            tm = np.array([s.time for s in filtered])
            n = filtered[0].mt.size
            model = _model_for_states(mod, tm) if len(filtered) > 1 else None
            if model is None:
                s = filtered[0]
                return [SmoothingState(s.time, s.mt, s.ct, s.at, s.rt)]
            packed = _pack_filtered(filtered)
            params = dict(V=np.eye(model.p), W=np.eye(n) if w is None else w)
            out = default_engine().smooth(model, params, packed, layout=SERIES_MAJOR,
                                          keep_init=True)
            _raise_status(int(out["status"][0]))
            return [SmoothingState(float(tm[r]), out["s"][0, r].copy(),
                                   _dlm.from_cm(out["S"][0, r], n, n).copy(),
                                   filtered[r].at, filtered[r].rt) for r in range(len(tm))]
        return run

    @staticmethod
    def ffbsDlm(mod: Dlm, ys: Sequence[Data], p: DlmParameters, *, z: Optional[np.ndarray] = None,
                rng: Optional[np.random.Generator] = None) -> List[SamplingState]:
        """``Smoothing.ffbsDlm`` (:173-180).  The N(0,1) draws are injected: ``z[row]`` are the
        n values used for theta[row]; with ``rng`` they are generated in the reference's
        draw order (last row first, MultivariateGaussianSvd.scala:19-22)."""
        model, params, times, y = _prep(mod, ys, p)
        rows, n = model.T + 1, model.n
        if z is None:
            rng = rng or np.random.default_rng()
            z = rng.standard_normal((rows, n))[::-1]
        z = np.ascontiguousarray(np.asarray(z, dtype=np.float64).reshape(1, rows, n))
        out = default_engine().ffbs(model, params, y, z, layout=SERIES_MAJOR,
                                    want_kf=("m", "C", "a", "R"))
        _raise_status(int(out["status"][0]))
        tm = _row_times(times, True)
        return [SamplingState(float(tm[r]), out["theta"][0, r].copy(), out["m"][0, r].copy(),
                              _dlm.from_cm(out["C"][0, r], n, n).copy(), out["a"][0, r].copy(),
                              _dlm.from_cm(out["R"][0, r], n, n).copy()) for r in range(rows)]


class SvdFilter:
    """``SvdFilter`` entry points (SvdFilter.scala)."""

    @staticmethod
    def filterDlm(mod: Dlm, ys: Sequence[Data], p: DlmParameters) -> List[SvdState]:
        """``SvdFilter.filterDlm`` (:158-161): T states; advance closure holds the raw W."""
        return SvdFilter._run(mod, ys, p, keep_init=False)

    @staticmethod
    def filter(mod: Dlm, ys: Sequence[Data], p: DlmParameters) -> List[SvdState]:
        """``SvdFilter(advanceState(p, mod.g)).filter`` (:112-119): T+1 states."""
        return SvdFilter._run(mod, ys, p, keep_init=True)

    @staticmethod
    def _run(mod, ys, p, keep_init):
        model, params, times, y = _prep(mod, ys, p)
        out = default_engine().svd_filter(model, params, y, layout=SERIES_MAJOR,
                                          keep_init=keep_init)
        _raise_status(int(out["status"][0]))
        n = model.n
        tm = _row_times(times, keep_init)
        return [SvdState(float(tm[r]), out["m"][0, r].copy(), out["dc"][0, r].copy(),
                         _dlm.from_cm(out["uc"][0, r], n, n).copy(), out["a"][0, r].copy(),
                         out["dr"][0, r].copy(), _dlm.from_cm(out["ur"][0, r], n, n).copy(),
                         out["f"][0, r].copy()) for r in range(len(tm))]


class SvdSampler:
    """``SvdSampler`` entry points (SvdSampler.scala)."""

    @staticmethod
    def ffbsDlm(mod: Dlm, ys: Sequence[Data], p: DlmParameters, *, z: Optional[np.ndarray] = None,
                rng: Optional[np.random.Generator] = None):
        """``SvdSampler.ffbsDlm`` (:79-82); returns (time, sample) per row."""
        model, params, times, y = _prep(mod, ys, p)
        rows, n = model.T + 1, model.n
        if z is None:
            rng = rng or np.random.default_rng()
            z = rng.standard_normal((rows, n))[::-1]
        z = np.ascontiguousarray(np.asarray(z, dtype=np.float64).reshape(1, rows, n))
        out = default_engine().ffbs(model, params, y, z, layout=SERIES_MAJOR, svd=True)
        _raise_status(int(out["status"][0]))
        tm = _row_times(times, True)
        return [(float(tm[r]), out["theta"][0, r].copy()) for r in range(rows)]


class GibbsSampling:
    """Sufficient statistics of ``GibbsSampling`` / ``GibbsWishart`` (Gibbs.scala:29-43,63-73)."""

    @staticmethod
    def sufficientStatistics(mod: Dlm, ys: Sequence[Data], theta: np.ndarray):
        times, y = _dlm.flatten_data(ys)
        model = Model.build(mod, times)
        yb = np.ascontiguousarray(y.reshape(1, model.T, model.p))
        th = np.ascontiguousarray(np.asarray(theta, dtype=np.float64).reshape(1, model.T + 1, model.n))
        out = default_engine().gibbs_stats(model, yb, th, layout=SERIES_MAJOR)
        return {k: v[0] for k, v in out.items()}

    @staticmethod
    def sample(mod: Dlm, priorV, priorW, initParams: DlmParameters, observations: Sequence[Data],
               n_iters: int, *, chains: int = 1, seed: int = 0, svd: bool = False):
        """``GibbsSampling.sample`` / ``sampleSvd`` (Gibbs.scala:153-217) -- or
        ``GibbsWishart.sample`` (GibbsWishart.scala:64-80) when ``priorW`` is an
        ``InverseWishart`` -- for ``chains`` independent chains on the same observations, every
        sweep on the GPU (FFBS + sufficient statistics + conjugate draws).  Returns a list (one
        entry per chain) of lists of ``DlmParameters`` (one per iteration)."""
        import torch
        from . import gibbs as _gibbs
        times, y = _dlm.flatten_data(observations)
        model = Model.build(mod, times)
        yb = torch.from_numpy(np.ascontiguousarray(np.repeat(y[None], chains, axis=0))).cuda()
        prior = dict(v_shape=priorV.shape, v_scale=priorV.scale)
        if isinstance(priorW, InverseWishart):
            prior.update(w_nu=priorW.nu, w_psi=np.asarray(priorW.psi, dtype=np.float64))
        else:
            prior.update(w_shape=priorW.shape, w_scale=priorW.scale)
        init = dict(V=initParams.v, W=initParams.w, m0=initParams.m0, C0=initParams.c0)
        res = _gibbs.sample(default_engine(), model, yb, prior, init, n_iters, seed=seed,
                            layout=SERIES_MAJOR, svd=svd)
        torch.cuda.synchronize()
        st = int(res["status"].max())
        _raise_status(st)
        V, W = res["V"].cpu().numpy(), res["W"].cpu().numpy()   # (iters, k, chains)
        n = model.n
        out = []
        for c in range(chains):
            out.append([DlmParameters(np.diag(V[i, :, c]),
                                      _dlm.from_cm(W[i, :, c], n, n) if W.shape[1] == n * n
                                      else np.diag(W[i, :, c]),
                                      initParams.m0, initParams.c0) for i in range(n_iters)])
        return out


@dataclass
class InverseGamma:
    """``InverseGamma(shape, scale)`` (InverseGamma.scala:6): prior on a variance."""
    shape: float
    scale: float

    @property
    def mean(self):
        return self.scale / (self.shape - 1)

    @property
    def variance(self):
        return (self.scale * self.scale) / ((self.shape - 1) * (self.shape - 1) * (self.shape - 2))


@dataclass
class InverseWishart:
    """``InverseWishart(nu, psi)`` (InverseWishart.scala:6): prior on a covariance matrix."""
    nu: float
    psi: np.ndarray


@dataclass
class SvParameters:
    """``SvParameters(phi, mu, sigmaEta)`` (StochasticVolatility.scala): AR(1) / OU state."""
    phi: float
    mu: float
    sigmaEta: float


@dataclass
class FilterState:
    """``FilterAr.FilterState`` (FilterAr.scala:9-13)."""
    time: float
    mt: float
    ct: float
    at: float
    rt: float


@dataclass
class SampleState:
    """``FilterAr.SampleState`` (FilterAr.scala:49-54)."""
    time: float
    sample: float
    mean: float
    cov: float
    at1: float
    rt1: float


class _ScalarFilter:
    ou = False

    @classmethod
    def _arrays(cls, ys, vs):
        if len(ys) == 0:
            raise ValueError("empty observation vector (ys.head on an empty Vector)")
        times = np.array([t for t, _ in ys], dtype=np.float64)
        y = np.array([[np.nan if o is None else float(o)] for _, o in ys], dtype=np.float64)
        return times, np.ascontiguousarray(y), np.ascontiguousarray(np.asarray(vs, dtype=np.float64))

    @classmethod
    def _row_times(cls, times):
        t0 = times[0] if cls.ou else times[0] - 1.0   # FilterOu.scala:37 / FilterAr.scala:42
        return np.concatenate([[t0], times])

    @classmethod
    def filterUnivariate(cls, ys, vs, p: SvParameters) -> List[FilterState]:
        """``filterUnivariate(ys: Vector[(Double, Option[Double])], vs, p)``."""
        times, y, v = cls._arrays(ys, vs)
        o = default_engine().ar_filter(dict(phi=p.phi, mu=p.mu, sigma_eta=p.sigmaEta), y, v,
                                       times=times, ou=cls.ou)
        tm = cls._row_times(times)
        return [FilterState(float(tm[r]), *(float(o[k][r, 0]) for k in ("m", "C", "a", "R")))
                for r in range(len(tm))]

    @classmethod
    def ffbs(cls, p: SvParameters, ys, vs, *, z: Optional[np.ndarray] = None, seed=None) -> List[SampleState]:
        """``ffbs(p, ys, vs)``: filter, then ``univariateSample``.  ``z``: injected N(0,1) values,
        one per row (row T is consumed first); drawn from ``seed`` when omitted."""
        times, y, v = cls._arrays(ys, vs)
        T = len(times)
        if z is None:
            z = np.random.default_rng(seed).standard_normal(T + 1)
        z = np.ascontiguousarray(np.asarray(z, dtype=np.float64).reshape(T + 1, 1))
        o = default_engine().ar_ffbs(dict(phi=p.phi, mu=p.mu, sigma_eta=p.sigmaEta), y, v, z,
                                     times=times, ou=cls.ou, want=("m", "C", "a", "R"))
        tm = cls._row_times(times)
        return [SampleState(float(tm[r]), float(o["theta"][r, 0]),
                            *(float(o[k][r, 0]) for k in ("m", "C", "a", "R"))) for r in range(T + 1)]


class FilterAr(_ScalarFilter):
    """``FilterAr`` (FilterAr.scala:8-83): scalar AR(1) state, unit time grid."""
    ou = False


class FilterOu(_ScalarFilter):
    """``FilterOu`` (FilterOu.scala:6-79): Ornstein-Uhlenbeck state, irregular times."""
    ou = True


@dataclass
class InverseGammaState:
    """``InverseGammaState(kfState, variance)`` (ConjugateFilter.scala:12)."""
    kfState: KfState
    variance: List[InverseGamma]


class ConjugateFilter:
    """``ConjugateFilter(prior, advState)`` (ConjugateFilter.scala:17-112), p = 1."""

    def __init__(self, prior: InverseGamma):
        self.prior = prior

    def filter(self, mod: Dlm, ys: Sequence[Data], p: DlmParameters) -> List[InverseGammaState]:
        times, y = _dlm.flatten_data(ys)
        model = Model.build(mod, times)
        yb = np.ascontiguousarray(y.reshape(1, model.T, model.p))
        o = default_engine().conjugate_filter(model, dict(W=p.w, m0=p.m0, C0=p.c0), self.prior.shape,
                                              self.prior.scale, yb, layout=SERIES_MAJOR,
                                              want=("m", "C", "a", "R", "f", "Q"))
        _raise_status(int(o["status"][0]))
        tm = _row_times(times, True)
        states = _kf_states(model, o, tm, True)
        return [InverseGammaState(s, [InverseGamma(float(o["shape"][0, r, 0]), float(o["scale"][0, r, 0]))])
                for r, s in enumerate(states)]
