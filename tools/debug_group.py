import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from bayesian_dlms_b200 import Engine, Model, TIME_MAJOR, SERIES_MAJOR
import helpers as H
mod, V, W, m0, C0 = H.seasonal13()
B, T = 9, 30
rng = np.random.default_rng(3)
y = np.stack([H.simulate(mod, V, W, m0, C0, np.arange(1, T + 1.0), rng, 0.1) for _ in range(B)])
model = Model.build(mod, T=T)
params = dict(V=V, W=W, m0=m0, C0=C0)
os.environ["BDLM_NO_GROUP_KERNEL"] = "1"
e_ref = Engine(0)
del os.environ["BDLM_NO_GROUP_KERNEL"]
e_grp = Engine(0)
for layout, yy in ((SERIES_MAJOR, y), (TIME_MAJOR, np.ascontiguousarray(y.transpose(1, 2, 0)))):
    yd = torch.from_numpy(yy).cuda()
    fr = e_ref.filter(model, params, yd, layout=layout); torch.cuda.synchronize()
    fg = e_grp.filter(model, params, yd, layout=layout); torch.cuda.synchronize()
    for k in ("m", "C", "a", "R", "f", "Q", "status"):
        a, b = fr[k].cpu().numpy(), fg[k].cpu().numpy()
        eq = (a == b) | (np.isnan(a) & np.isnan(b))
        print(layout, "filter", k, "equal" if eq.all() else f"DIFF {np.sum(~eq)}/{a.size}")
    sr = e_ref.smooth(model, params, fr, layout=layout); torch.cuda.synchronize()
    sg = e_grp.smooth(model, params, fg, layout=layout); torch.cuda.synchronize()
    sg2 = e_ref.smooth(model, params, fg, layout=layout); torch.cuda.synchronize()
    for k in ("s", "S"):
        print(layout, "smooth", k, "grp-engine:", torch.equal(sr[k], sg[k]), " ref-engine on grp outputs:", torch.equal(sr[k], sg2[k]))
print("---- per-series parameters")
from bayesian_dlms_b200 import dlm
sc = np.exp(rng.uniform(np.log(0.5), np.log(2.0), size=(B, 2)))
Vs = np.stack([dlm.cm(V * s[0]) for s in sc]); Ws = np.stack([dlm.cm(W * s[1]) for s in sc]); m0s = rng.standard_normal((B, 13))
for layout, yy in ((SERIES_MAJOR, y), (TIME_MAJOR, np.ascontiguousarray(y.transpose(1, 2, 0)))):
    tr = (lambda a: np.ascontiguousarray(a.T)) if layout == TIME_MAJOR else (lambda a: a)
    pp = dict(V=torch.from_numpy(tr(Vs)).cuda(), W=torch.from_numpy(tr(Ws)).cuda(), m0=torch.from_numpy(tr(m0s)).cuda(), C0=C0, per_series=("V", "W", "m0"))
    yd = torch.from_numpy(yy).cuda()
    fr = e_ref.filter(model, pp, yd, layout=layout); torch.cuda.synchronize()
    fg = e_grp.filter(model, pp, yd, layout=layout); torch.cuda.synchronize()
    for k in ("m", "C", "a", "R", "f", "Q", "status"):
        print(layout, "filter", k, torch.equal(fr[k].nan_to_num(), fg[k].nan_to_num()))
    sr = e_ref.smooth(model, pp, fr, layout=layout); torch.cuda.synchronize()
    sg = e_grp.smooth(model, pp, fg, layout=layout); torch.cuda.synchronize()
    sg3 = e_grp.smooth(model, pp, fg, layout=layout); torch.cuda.synchronize()
    print(layout, "smooth", torch.equal(sr["s"], sg["s"]), torch.equal(sr["S"], sg["S"]), "repeat:", torch.equal(sr["s"], sg3["s"]))
    fg2 = e_grp.filter(model, pp, yd, layout=layout)
    sg4 = e_grp.smooth(model, pp, fg2, layout=layout); torch.cuda.synchronize()
    print(layout, "filter->smooth back to back, no sync:", torch.equal(sr["s"], sg4["s"]), torch.equal(sr["S"], sg4["S"]))
print("---- which side is wrong?")
layout = SERIES_MAJOR
pp = dict(V=torch.from_numpy(Vs).cuda(), W=torch.from_numpy(Ws).cuda(), m0=torch.from_numpy(m0s).cuda(), C0=C0, per_series=("V", "W", "m0"))
yd = torch.from_numpy(y).cuda()
fr = e_ref.filter(model, pp, yd, layout=layout); torch.cuda.synchronize()
sr = e_ref.smooth(model, pp, fr, layout=layout); torch.cuda.synchronize()
for trial in range(3):
    fg2 = e_grp.filter(model, pp, yd, layout=layout)
    sg4 = e_grp.smooth(model, pp, fg2, layout=layout); torch.cuda.synchronize()
    print("trial", trial, "filter outputs ok:", all(torch.equal(fr[k], fg2[k]) for k in ("m", "C", "a", "R")),
          "smooth ok:", torch.equal(sr["s"], sg4["s"]), "status", fg2["status"].tolist(), sg4["status"].tolist())
    bad = (sr["s"] != sg4["s"]).nonzero()
    print("   first bad idx", bad[:3].tolist(), "n bad", len(bad))
# same with shared params
pp2 = dict(V=V, W=W, m0=m0, C0=C0)
fr = e_ref.filter(model, pp2, yd, layout=layout); torch.cuda.synchronize()
sr = e_ref.smooth(model, pp2, fr, layout=layout); torch.cuda.synchronize()
fg2 = e_grp.filter(model, pp2, yd, layout=layout)
sg4 = e_grp.smooth(model, pp2, fg2, layout=layout); torch.cuda.synchronize()
print("shared params back-to-back: smooth ok:", torch.equal(sr["s"], sg4["s"]))
