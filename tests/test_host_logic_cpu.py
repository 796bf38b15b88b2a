"""Host-side model assembly (CPU): flattening the Dlm closures for a time grid."""
import numpy as np

from bayesian_dlms_b200 import dlm
from bayesian_dlms_b200.batch import Model


def _brute(mod, times, t_init=None):
    t0 = times.min() - 1.0 if t_init is None else t_init
    prev = np.concatenate([[t0], times[:-1]])
    F = np.stack([dlm.cm(mod.f(float(t))) for t in times])
    G = np.stack([dlm.cm(mod.g(float(dt))) for dt in times - prev])
    return F, G


def test_materialise_matches_per_step_evaluation():
    rng = np.random.default_rng(0)
    times = np.cumsum(rng.choice([0.5, 1.0, 2.0], 40))
    for mod in (dlm.polynomial(2), dlm.polynomial(1) + dlm.seasonal(24, 3),
                dlm.polynomial(1) * dlm.polynomial(1),
                dlm.regression([np.array([float(i)]) for i in range(60)])):
        tt = np.arange(1.0, 41.0) if not getattr(mod, "f_const", False) else times
        F, f_tv, G, g_tv, n, p = dlm.materialise(mod, tt)
        Fb, Gb = _brute(mod, tt)
        assert np.array_equal(F if f_tv else np.tile(F, (len(tt), 1)), Fb)
        assert np.array_equal(G if g_tv else np.tile(G, (len(tt), 1)), Gb)
    assert dlm.polynomial(3).f_const and (dlm.polynomial(1) + dlm.seasonal(12, 2)).f_const
    assert not dlm.regression([np.zeros(1)] * 3).f_const
    assert not (dlm.polynomial(1) * dlm.regression([np.zeros(1)] * 3)).f_const


def test_regular_grid_is_detected_and_long_grids_are_cheap():
    m = Model.build(dlm.polynomial(2), T=1 << 20)       # one f and one g evaluation
    assert m.times is None and not m.f_tv and not m.g_tv and m.T == 1 << 20
    m = Model.build(dlm.polynomial(1) + dlm.seasonal(24, 2), times=np.arange(1.0, 11.0))
    assert m.times is None and not m.g_tv
    m = Model.build(dlm.polynomial(1) + dlm.seasonal(24, 2), times=np.array([1.0, 2.0, 4.0]))
    assert m.times is not None and m.g_tv and m.G.size == 3 * 25


def test_resume_time_sets_the_first_increment():
    mod = dlm.polynomial(1) + dlm.seasonal(24, 2)
    times = np.array([10.0, 11.0, 13.0])
    F, f_tv, G, g_tv, n, p = dlm.materialise(mod, times, t_init=9.5)
    _, Gb = _brute(mod, times, 9.5)
    assert g_tv and np.array_equal(G, Gb)
    m = Model.build(mod, times, t_init=9.5)
    assert m.t_init is not None and m.t_init[0] == 9.5 and m.times is not None
    # forecast grid: first increment 0 (oneStepPrediction on the state itself)
    F, f_tv, G, g_tv, n, p = dlm.materialise(mod, 20.0 + np.arange(4.0), t_init=20.0)
    assert np.array_equal(G[0], dlm.cm(mod.g(0.0))) and np.array_equal(G[1], dlm.cm(mod.g(1.0)))
