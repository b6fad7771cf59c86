"""Multi-rank host logic on CPU (gloo, world_size 2): slab partition of the global env index,
seed assignment independent of the number of ranks, and the episode-statistics all-reduce."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from balatro_gym_b200 import dist as bdist


def test_slab_partition_covers_everything():
    for total in (1, 7, 1000, 1 << 20, (1 << 20) + 3):
        for ws in (1, 2, 3, 8):
            spans = [bdist.slab(total, r, ws) for r in range(ws)]
            assert spans[0][0] == 0
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert spans[-1][0] + spans[-1][1] == total
            sizes = [c for _, c in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank))
    r, lr, w = bdist.init_process_group("gloo")
    assert (r, w) == (rank, ws)
    start, count = bdist.slab(total, rank, ws)
    # seeds are a function of the GLOBAL env index (what BalatroVecEnv.default_seeds computes)
    seeds = (torch.arange(count, dtype=torch.int64) + (1 + start)) % (2 ** 32)
    # per-slab statistics vector: [episodes, sum_return, sum_length, steps, sum_reward, 0, 0, 0]
    stats = torch.tensor([count, float(seeds.sum()), 2.0 * count, 10.0 * count, -1.0 * count, 0, 0, 0], dtype=torch.float64)
    bdist.allreduce_stats(stats)
    t = bdist.max_over_ranks(float(rank + 1))
    # policy-side gradient averaging of the data-parallel PPO update (rollout.ppo_update)
    lin = torch.nn.Linear(3, 2)
    for p in lin.parameters():
        p.grad = torch.full_like(p, float(rank + 1))
    bdist.allreduce_mean_grads(list(lin.parameters()))
    assert all(bool((p.grad == 1.5).all()) for p in lin.parameters())
    bdist.barrier()
    q.put((rank, start, count, stats.tolist(), t, seeds[:3].tolist()))
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_and_seeds():
    total, ws = 1001, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, total, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, c0, st0, t0, sd0), (r1, s1, c1, st1, t1, sd1) = res
    assert s0 == 0 and s0 + c0 == s1 and s1 + c1 == total
    assert st0 == st1                                   # both ranks hold the reduced vector
    assert st0[0] == total and st0[2] == 2.0 * total
    assert st0[1] == float(sum(range(1, total + 1)))    # every global env index seeded exactly once
    assert t0 == t1 == 2.0                              # max over ranks
    assert sd0 == [1, 2, 3] and sd1 == [s1 + 1, s1 + 2, s1 + 3]
