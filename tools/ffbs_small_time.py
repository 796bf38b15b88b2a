"""Time FFBS for the config-2 model (polynomial(2), n = 2, p = 1): one thread per chain."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesian_dlms_b200 import Engine, Model, TIME_MAJOR, dlm
ap = argparse.ArgumentParser(); ap.add_argument("--B", type=int, default=200_000); ap.add_argument("--T", type=int, default=1000)
ap.add_argument("--n", type=int, default=2)
a = ap.parse_args()
eng = Engine(0)
B, T, n = a.B, a.T, a.n
g = torch.Generator(device="cuda").manual_seed(1)
y = torch.randn((T, 1, B), generator=g, device="cuda", dtype=torch.float64).cumsum(0)
z = torch.randn((T + 1, n, B), generator=g, device="cuda", dtype=torch.float64)
params = dict(V=[[3.0]], W=np.diag(np.linspace(2.0, 1.0, n)) if n > 1 else [[3.0]], m0=np.zeros(n), C0=100.0 * np.eye(n))
model = Model.build(dlm.polynomial(n), T=T)
for label, zz in (("injected z", z), ("Philox", None)):
    eng.ffbs(model, params, y, zz, stats=True); torch.cuda.synchronize()
    ms = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = eng.ffbs(model, params, y, zz, stats=True); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    med = float(np.median(ms))
    print(f"ffbs n={n} B={B} T={T} ({label}): {med:.2f} ms  {B / med * 1e3:.0f} draws/s  {B * T / med / 1e6:.3f} G steps/s "
          f"status={int(out['status'].max())}")
