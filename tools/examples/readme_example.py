"""The batched-use example of README.md as a runnable script (needs a B200)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from bayesian_dlms_b200 import Engine, Model, TIME_MAJOR, dlm, gibbs
eng = Engine(0)
mod = dlm.polynomial(2)
model = Model.build(mod, T=1000)
y = torch.randn(1000, 1, 10_000, device="cuda", dtype=torch.float64).cumsum(0)
params = dict(V=[[3.0]], W=np.diag([2.0, 1.0]), m0=np.zeros(2), C0=100.0 * np.eye(2))
out = eng.filter_smooth(model, params, y, layout=TIME_MAJOR)
ll = eng.loglik(model, params, y)
chains = gibbs.sample(eng, model, y, dict(v_shape=3.0, v_scale=4.0, w_shape=3.0, w_scale=6.0),
                      params, n_iters=5, seed=1)
torch.cuda.synchronize()
print(sorted(out), ll["transition"].shape, chains["V"].shape, chains["W"].shape, int(chains["status"].max()))
