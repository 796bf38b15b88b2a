"""Time the time-sharded scan of ONE series (T = 2^24) from ONE process driving all visible GPUs
through bdlm_comm_scan_filter_smooth: NCCL all-gathers vs peer-mailbox exchange.

    python tools/comm_scan_time.py [--logT 24] [--reps 10]

Device time = per-device CUDA events on the stream each context runs on (max over devices);
host time = wall clock around enqueue + sync from the single host thread."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesian_dlms_b200 import Model, dlm  # noqa: E402
from bayesian_dlms_b200.comm import Comm  # noqa: E402
from bayesian_dlms_b200.sharding import shard_range  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--logT", type=int, default=24)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--only", choices=("nccl", "peer"), default=None)
a = ap.parse_args()
world = torch.cuda.device_count()
T = 1 << a.logT
params = dict(V=[[3.0]], W=np.diag([2.0, 1.0]), m0=np.zeros(2), C0=100.0 * np.eye(2))
g = torch.Generator(device="cuda:0").manual_seed(20260105)
yfull = torch.randn(T, generator=g, device="cuda:0", dtype=torch.float64).cumsum(0) * 0.1
for peer in (False, True):
    if a.only is not None and peer != (a.only == "peer"):
        continue
    if peer:
        os.environ.pop("BDLM_COMM_NO_PEER", None)
    else:
        os.environ["BDLM_COMM_NO_PEER"] = "1"
    comm = Comm.single_process(list(range(world)))
    streams = []
    for r, e in enumerate(comm.engines):     # contexts on torch streams so torch events see them
        with torch.cuda.device(r):
            s = torch.cuda.Stream(device=r)
            streams.append(s)
            e.ctx.set_stream(s.cuda_stream)
    models, chunks = [], []
    for r in range(world):
        lo, hi = shard_range(T, r, world)
        models.append(Model.build(dlm.polynomial(2), T=hi - lo))
        chunks.append(yfull[lo:hi].to(f"cuda:{r}").contiguous())
    for r in range(world):
        torch.cuda.synchronize(r)
    # scan_setup binds engines to torch's CURRENT stream of each device: make it ours
    ctxs = [torch.cuda.stream(s) for s in streams]
    for c in ctxs:
        c.__enter__()
    h = comm.scan_setup(models, params, chunks)
    for _ in range(3):
        comm.scan_run(h)
    comm.sync()
    dev_ms, host_ms = [], []
    for _ in range(a.reps):
        comm.sync()
        evs = []
        for r, s in enumerate(streams):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            evs.append((e0, e1))
        w0 = time.perf_counter()
        comm.scan_run(h)
        w1 = time.perf_counter()
        for (e0, e1), s in zip(evs, streams):
            e1.record(s)
        comm.sync()
        w2 = time.perf_counter()
        dev_ms.append(max(e0.elapsed_time(e1) for e0, e1 in evs))
        host_ms.append(((w1 - w0) * 1e3, (w2 - w0) * 1e3))
    for c in reversed(ctxs):
        c.__exit__(None, None, None)
    print(f"world={world} peer_mailboxes={comm.uses_peer_exchange} T=2^{a.logT}: device {np.median(dev_ms):.3f} ms "
          f"(max over devices, events), host enqueue {np.median([x[0] for x in host_ms]):.3f} ms, "
          f"enqueue+sync {np.median([x[1] for x in host_ms]):.3f} ms, status "
          f"{max(int(s.item()) for s in h['status'])}")
    comm.close()
