"""Time KalmanFilter.likelihood batched over series (config-2 model, per-series V, W)."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesian_dlms_b200 import Engine, Model, TIME_MAJOR, dlm
ap = argparse.ArgumentParser(); ap.add_argument("--B", type=int, default=100_000); ap.add_argument("--T", type=int, default=1000)
a = ap.parse_args()
eng = Engine(0)
B, T = a.B, a.T
g = torch.Generator(device="cuda").manual_seed(1)
y = torch.randn((T, 1, B), generator=g, device="cuda", dtype=torch.float64).cumsum(0)
sc = torch.exp(torch.rand((2, B), generator=g, device="cuda", dtype=torch.float64) * 1.386 - 0.693)
params = dict(V=(3.0 * sc[0:1]).contiguous(),
              W=(torch.tensor([2.0, 0.0, 0.0, 1.0], device="cuda", dtype=torch.float64)[:, None] * sc[1:2]).contiguous(),
              m0=np.zeros(2), C0=100.0 * np.eye(2), per_series=("V", "W"))
model = Model.build(dlm.polynomial(2), T=T)
eng.loglik(model, params, y); torch.cuda.synchronize()
ms = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = eng.loglik(model, params, y); e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
med = float(np.median(ms))
print(f"loglik B={B} T={T}: {med:.2f} ms  {B * T / med / 1e6:.2f} G series-steps/s  status={int(out['status'].max())} "
      f"sum_tr={float(out['transition'].sum()):.6e} sum_in={float(out['innovations'].sum()):.6e}")
