#!/usr/bin/env python
"""Probe: one slab of 2^20 envs on one stream vs S slabs of 2^20 / S envs, each on its own stream (no cross-slab
ordering: the phases of different slabs interleave on the GPU, so one slab's latency-bound tail overlaps the others'
bandwidth-bound main pass).  Same envs, same seeds (seed = f(global env index)), same total work.

    python tools/exp/slabs_probe.py [n_envs] [steps]
"""
import sys
import time

import torch

sys.path.insert(0, ".")
import balatro_gym_b200 as b  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
K = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)


def run(S, fused=False, burn=150, graph=False):
    m = n // S
    envs = [b.BalatroVecEnv(m, device=dev, seed=1, autoreset=True, env_offset=i * m, generator="c4") for i in range(S)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
    for e in envs:
        e.reset()
    torch.cuda.synchronize(dev)

    def step_all():
        for e, st in zip(envs, streams):
            with torch.cuda.stream(st):
                if fused:
                    e.step(random_policy=True, want_info=False)
                else:
                    e.sample_actions(seed=2024)
                    e.step(e.actions, want_info=False)

    for _ in range(burn):
        step_all()
    torch.cuda.synchronize(dev)
    if graph:      # one CUDA graph per slab and step, replayed on the slab's stream
        replays = []
        for e, st in zip(envs, streams):
            with torch.cuda.stream(st):
                replays.append(e.graphed_rollout_step("fused" if fused else "sampler", seed=2024))
        torch.cuda.synchronize(dev)

        def step_all():  # noqa: F811
            for r, st in zip(replays, streams):
                with torch.cuda.stream(st):
                    r()
        for _ in range(10):
            step_all()
        torch.cuda.synchronize(dev)
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(S)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(S)]
    t0 = time.perf_counter()
    for st, e0 in zip(streams, ev0):
        e0.record(st)
    for _ in range(K):
        step_all()
    for st, e1 in zip(streams, ev1):
        e1.record(st)
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0
    ms = max(ev0[0].elapsed_time(e1) for e1 in ev1)
    ph = torch.cat([e.state_field("phase").long() for e in envs])
    chk = int(torch.cat([e.state_field("chips_scored") for e in envs]).sum())
    print(f"slabs {S} fused {int(fused)} graph {int(graph)}: {ms / K * 1e3:8.1f} us/step  {n * K / ms * 1e3:.3e} env-steps/s  (wall {wall * 1e3 / K * 1e3:.1f} us/step)"
          f"  play-phase frac {float((ph == 0).double().mean()):.3f}  chips checksum {chk}", flush=True)
    del envs


for graph in (True,):
    for fused in (False, True):
        for S in (1, 2, 3, 4, 8):
            run(S, fused, graph=graph)
