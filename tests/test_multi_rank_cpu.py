"""World-size-2 gloo test of the N > 1 host logic (runs on CPU): series are sharded across
ranks with no data-path collective; only per-rank scalars (log-likelihood sum, pooled Gibbs
statistics) are sum-reduced.  The per-series numbers come from the oracle here -- the point is
the sharding / reduction plumbing, which is what bench.py uses under torchrun on GPUs."""
import os
import socket

import numpy as np
import pytest

from bayesian_dlms_b200.sharding import shard_range, wave_aligned_slabs


def test_shard_ranges_partition_the_batch():
    for total in (0, 1, 7, 8, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            r = [shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            for a, b in zip(r, r[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_wave_aligned_slabs():
    s = wave_aligned_slabs(0, 1_000_000, 94_720, 5)
    assert s[0] == (0, 473_600) and s[-1][1] == 1_000_000
    assert all(b - a == 473_600 for a, b in s[:-1])
    assert sum(b - a for a, b in s) == 1_000_000


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, q):
    import torch
    import torch.distributed as dist
    import oracle
    from bayesian_dlms_b200 import dlm
    from bayesian_dlms_b200.sharding import reduce_max, reduce_sum
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(B, rank, world)
        rng = np.random.default_rng(5)
        y = rng.standard_normal((B, 30, 1)).cumsum(axis=1)  # every rank builds the same data
        mod = dlm.polynomial(2)
        times = np.arange(1, 31.0)
        F, _, G, _, n, p = dlm.materialise(mod, times)
        ll = 0.0
        for b in range(lo, hi):  # each rank only touches its own block
            ll += oracle.loglik(n, p, F, G, [3.0], dlm.cm(np.diag([2.0, 1.0])), np.zeros(2),
                                dlm.cm(100 * np.eye(2)), times, y[b])["innovations"]
        t = torch.tensor([ll, float(hi - lo)], dtype=torch.float64)
        reduce_sum(t)
        tm = torch.tensor([float(rank)], dtype=torch.float64)
        reduce_max(tm)
        q.put((rank, float(t[0]), float(t[1]), float(tm[0]), lo, hi))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shard_and_reduce():
    import torch.multiprocessing as mp
    import oracle
    from bayesian_dlms_b200 import dlm
    oracle.build()
    B, world = 21, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process answer
    rng = np.random.default_rng(5)
    y = rng.standard_normal((B, 30, 1)).cumsum(axis=1)
    times = np.arange(1, 31.0)
    F, _, G, _, n, p_ = dlm.materialise(dlm.polynomial(2), times)
    full = sum(oracle.loglik(n, p_, F, G, [3.0], dlm.cm(np.diag([2.0, 1.0])), np.zeros(2),
                             dlm.cm(100 * np.eye(2)), times, y[b])["innovations"] for b in range(B))
    assert res[0][4:] == (0, 11) and res[1][4:] == (11, 21)
    for rank, ll, cnt, mx, lo, hi in res:
        assert cnt == B and mx == world - 1
        assert abs(ll - full) <= 1e-12 * abs(full)
