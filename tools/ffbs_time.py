"""Time the warp-per-series kernels on the config-3 / config-4 shapes (FFBS draws/s).

    python tools/ffbs_time.py --config 3 [--B 4096] [--T 2000]
    python tools/ffbs_time.py --config 4 [--B 65536] [--T 1000]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from bayesian_dlms_b200 import Engine, Model, SERIES_MAJOR  # noqa: E402
import helpers as H  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=3)
ap.add_argument("--B", type=int, default=0)
ap.add_argument("--T", type=int, default=0)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--op", default="ffbs", choices=["ffbs", "svd_ffbs", "filter", "filter_smooth", "svd_filter", "loglik"])
a = ap.parse_args()
dev = torch.device("cuda", 0)
eng = Engine(0)
if a.config == 3:
    mod, V, W, m0, C0 = H.seasonal13()
    B, T, miss = a.B or 4096, a.T or 2000, 0.1
else:
    mod, V, W, m0, C0 = H.correlated8()
    B, T, miss = a.B or 65536, a.T or 1000, 0.0
n, p = len(m0), V.shape[0]
g = torch.Generator(device=dev).manual_seed(7)
y = torch.randn((B, T, p), generator=g, device=dev, dtype=torch.float64) * 2.0
if miss:
    y[torch.rand((B, T, p), generator=g, device=dev) < miss] = float("nan")
z = torch.randn((B, T + 1, n), generator=g, device=dev, dtype=torch.float64)
model = Model.build(mod, T=T)
params = dict(V=V, W=W, m0=m0, C0=C0)


def run():
    if a.op == "ffbs":
        return eng.ffbs(model, params, y, z, layout=SERIES_MAJOR, stats=True)
    if a.op == "svd_ffbs":
        return eng.ffbs(model, params, y, z, layout=SERIES_MAJOR, stats=True, svd=True)
    if a.op == "filter":
        return eng.filter(model, params, y, layout=SERIES_MAJOR, want=("m", "C"))
    if a.op == "filter_smooth":
        return eng.filter_smooth(model, params, y, layout=SERIES_MAJOR, want=("s", "S"))
    if a.op == "svd_filter":
        return eng.svd_filter(model, params, y, layout=SERIES_MAJOR, want=("m", "dc", "uc"))
    return eng.loglik(model, params, y, layout=SERIES_MAJOR)


out = run()
torch.cuda.synchronize()
ms = []
for _ in range(a.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = run()
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
med = float(np.median(ms))
st = int(out["status"].max()) if "status" in out else -1
print(f"config={a.config} op={a.op} n={n} p={p} B={B} T={T} median={med:.1f} ms  "
      f"{B / med * 1e3:.0f} draws(series)/s  {B * (T + 1) / med / 1e3:.2f} M steps/s status_max={st}")
