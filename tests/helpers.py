"""Shared synthetic-workload builders for the parity tests."""
import csv
import json
import os

import numpy as np

from bayesian_dlms_b200 import dlm

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def read_csv(name):
    with open(os.path.join(GOLD, name)) as fh:
        rows = list(csv.reader(fh))[1:]
    return rows


def first_order_golden():
    rows = read_csv("first_order_dlm.csv")
    times = np.array([float(r[0]) for r in rows])
    y = np.array([float(r[1]) for r in rows]).reshape(-1, 1)
    filt = read_csv("first_order_dlm_filtered.csv")
    sm = read_csv("first_order_dlm_smoothed.csv")
    g = dict(
        m=np.array([float(r[1]) for r in filt]), C=np.array([float(r[2]) for r in filt]),
        f=np.array([float(r[3]) for r in filt[1:]]), Q=np.array([float(r[4]) for r in filt[1:]]),
        s=np.array([float(r[1]) for r in sm]), S=np.array([float(r[2]) for r in sm]),
        time=np.array([float(r[0]) for r in filt]))
    return times, y, g


def kat():
    with open(os.path.join(GOLD, "kat.json")) as fh:
        return json.load(fh)


def spd(rng, n, scale=1.0):
    A = rng.standard_normal((n, n))
    return scale * (A @ A.T / n + 0.5 * np.eye(n))


def simulate(mod, V, W, m0, C0, times, rng, missing=0.0):
    """Dlm.simStep generative model (Dlm.scala:245-282), numpy RNG."""
    times = np.asarray(times, float)
    n = len(m0)
    p = V.shape[0]
    x = rng.multivariate_normal(m0, C0)
    prev = times.min() - 1.0
    ys = np.empty((times.size, p))
    for i, t in enumerate(times):
        dt = t - prev
        x = mod.g(dt) @ x + (rng.multivariate_normal(np.zeros(n), W * dt) if dt > 0 else 0)
        ys[i] = mod.f(t).T @ x + rng.multivariate_normal(np.zeros(p), V)
        prev = t
    if missing > 0:
        ys[rng.random(ys.shape) < missing] = np.nan
    return ys


def seasonal13():
    """Config 3 model: polynomial(1) |+| seasonal(24, 6), n = 13, p = 1."""
    mod = dlm.polynomial(1) + dlm.seasonal(24, 6)
    pat = [0.2, 0.4, 0.5, 0.2, 0.1, 0.4]
    W = np.diag([0.01] + pat + pat)
    return mod, np.array([[1.0]]), W, np.zeros(13), np.eye(13)


def correlated8():
    """Config 4 model: 8-fold outer sum of polynomial(1), n = p = 8, full W."""
    mod = dlm.polynomial(1)
    for _ in range(7):
        mod = mod * dlm.polynomial(1)
    V = np.diag([1.0, 4.0] * 4)
    W = np.diag([0.75, 1.25] * 4) + 0.5 * (np.eye(8, k=1) + np.eye(8, k=-1))
    return mod, V, W, np.zeros(8), np.eye(8)


def second_order():
    """Config 2 model: polynomial(2), V = 3, W = diag(2, 1), C0 = 100 I."""
    return (dlm.polynomial(2), np.array([[3.0]]), np.diag([2.0, 1.0]), np.zeros(2),
            100.0 * np.eye(2))


def rel_err(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    den = np.maximum(np.abs(b), 1e-300)
    scale = np.max(np.abs(b)) if b.size else 1.0
    # relative to the entry, floored at 1e-6 of the array scale so exact zeros compare sanely
    den = np.maximum(den, 1e-6 * max(scale, 1e-300))
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0
