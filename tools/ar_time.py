"""Time the scalar AR(1) / OU filter + backward sampler (next row f3) on the bench shape.

    python tools/ar_time.py [--B 1000000] [--T 1000] [--ou] [--reps 3]
"""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesian_dlms_b200 import Engine, TIME_MAJOR
ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=1_000_000); ap.add_argument("--T", type=int, default=1000)
ap.add_argument("--ou", action="store_true"); ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
eng = Engine(0)
g = torch.Generator(device=dev).manual_seed(6)
r = lambda *s: torch.rand(s, generator=g, device=dev, dtype=torch.float64)  # noqa: E731
B, T = a.B, a.T
sv = dict(phi=0.5 + 0.45 * r(B), mu=r(B) - 0.5, sigma_eta=0.1 + r(B))
y = torch.randn((T, B), generator=g, device=dev, dtype=torch.float64)
v = 0.5 + r(T, B)
z = torch.randn((T + 1, B), generator=g, device=dev, dtype=torch.float64)
times = np.cumsum(np.random.default_rng(0).uniform(0.5, 1.5, T)) if a.ou else None
for name, fn, byt in (("filter", lambda: eng.ar_filter(sv, y, v, times=times, ou=a.ou, layout=TIME_MAJOR), 8 * 6),
                      ("ffbs", lambda: eng.ar_ffbs(sv, y, v, z, times=times, ou=a.ou, layout=TIME_MAJOR), 8 * 8)):
    fn(); torch.cuda.synchronize()
    ms = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
        del out
    med = float(np.median(ms))
    print(f"{'ou' if a.ou else 'ar1'} {name}: B={B} T={T} median={med:.2f} ms  {B * T / med / 1e6:.2f} G series-steps/s  "
          f"{B * T * byt / med / 1e6:.0f} GB/s algorithmic ({byt} B/step)")
