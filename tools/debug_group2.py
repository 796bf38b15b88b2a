import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from bayesian_dlms_b200 import Engine, Model, SERIES_MAJOR
import helpers as H
mod, V, W, m0, C0 = H.seasonal13()
B, T = 9, 30
rng = np.random.default_rng(3)
y = np.stack([H.simulate(mod, V, W, m0, C0, np.arange(1, T + 1.0), rng, 0.1) for _ in range(B)])
model = Model.build(mod, T=T)
pp = dict(V=V, W=W, m0=m0, C0=C0)
eng = Engine(0)
yd = torch.from_numpy(y).cuda()
f = eng.filter(model, pp, yd, layout=SERIES_MAJOR); torch.cuda.synchronize()
s_ok = eng.smooth(model, pp, f, layout=SERIES_MAJOR); torch.cuda.synchronize()
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
f2 = eng.filter(model, pp, yd, layout=SERIES_MAJOR)
if mode == "bdlm_sync": eng.sync()
if mode == "sleep":
    import time; time.sleep(0.5)
s2 = eng.smooth(model, pp, f2, layout=SERIES_MAJOR); torch.cuda.synchronize()
print(mode, "CUDA_LAUNCH_BLOCKING=", os.environ.get("CUDA_LAUNCH_BLOCKING"), "smooth ok:", torch.equal(s_ok["s"], s2["s"]),
      "filter same:", torch.equal(f["m"], f2["m"]))
# third: smooth twice on f2 after sync
s3 = eng.smooth(model, pp, f2, layout=SERIES_MAJOR); torch.cuda.synchronize()
print("   smooth again on the same f2 after sync:", torch.equal(s_ok["s"], s3["s"]))
