"""Time the parallel-in-time filter+smoother on one long series (config 5 shape).

    python tools/scan_time.py --cases 2:24,1:24,2:20     # n:log2(T) pairs, one process
"""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesian_dlms_b200 import Engine, Model, TIME_MAJOR, dlm
from bayesian_dlms_b200.scan import scan_filter_smooth
ap = argparse.ArgumentParser()
ap.add_argument("--cases", default="2:24")
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
eng = Engine(0)
for case in a.cases.split(","):
    n, logT = (int(v) for v in case.split(":"))
    T = 1 << logT
    mod = dlm.polynomial(n)
    params = dict(V=[[3.0]], W=np.diag(np.linspace(2.0, 1.0, n)) if n > 1 else [[3.0]], m0=np.zeros(n), C0=100.0 * np.eye(n))
    y = torch.randn(T, device="cuda", dtype=torch.float64).cumsum(0) * 0.1 + torch.randn(T, device="cuda", dtype=torch.float64)
    model = Model.build(mod, T=T)
    for _ in range(2):
        out = scan_filter_smooth(eng, model, params, y)
    torch.cuda.synchronize()
    ms = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = scan_filter_smooth(eng, model, params, y); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    med = float(np.median(ms))
    byt = 8 * (1 + 2 * n + 2 * n * n + 2 + 2 * (n + n * n))  # y, KfState, (m,C) re-read, (s,S)
    print(f"scan n={n} T=2^{logT} median={med:.3f} ms  {T / med / 1e6:.2f} G steps/s  {T * byt / med / 1e6:.0f} GB/s algorithmic status={int(out['status'][0])}")
    # sequential single-thread GPU kernel for comparison (one series = one thread)
    if logT <= 20:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.filter_smooth(model, params, y.reshape(-1, 1, 1), layout=TIME_MAJOR, textbook=True); e1.record(); torch.cuda.synchronize()
        print(f"sequential kernel (1 thread): {e0.elapsed_time(e1):.1f} ms")
    del out, y
