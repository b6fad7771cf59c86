#!/usr/bin/env python
"""Reference-vs-CUDA differential run on the GPU box, with fresh seeds (not the committed golden traces):
the unmodified reference (byte-compiled copy under oracle/_ref) plays episodes with every RNG draw tapped — in
worker processes, one per host core — then the CUDA env replays the same actions and draws through the C-ABI and
every step is compared: state record, observation record, reward, termination, score breakdown, error flag.

    python tools/lockstep_cuda.py --steps 1000000 --seed0 500001
Writes gpurun_out/lockstep_cuda.txt.  tests/test_gpu_lockstep.py runs the same thing under pytest.
"""
import argparse
import multiprocessing as mp
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

CONFIGS = ("c1", "c3", "c4", "c4x")
MAX_STEPS = {"c1": 600, "c3": 600, "c4": 600, "c4x": 120}


def _record_chunk(job):
    """Worker: record one batch of reference episodes, park it in a temporary .npz, return (path, steps, raised)."""
    cfg, episodes, seed0, out_dir = job
    from make_golden import record
    tr = record(cfg, episodes, seed0, MAX_STEPS[cfg])
    path = os.path.join(out_dir, f"{cfg}_{seed0}.npz")
    np.savez(path, **{k: (v.view(np.uint8) if v.dtype.fields else v) for k, v in tr.items()})
    return path, cfg, int(tr["length"].sum()), int(tr["exc"].sum())


def _load(path):
    from balatro_gym_b200 import layout as L
    z = np.load(path)
    out = {k: z[k] for k in z.files}
    T, E = out["action"].shape
    out["draws"] = out["draws"].reshape(T, -1).view(L.DRAWS_DTYPE).reshape(T, E)
    out["state"] = out["state"].reshape(T, -1).view(L.STATE_DTYPE).reshape(T, E)
    out["obs"] = out["obs"].reshape(T, -1).view(L.OBS_DTYPE).reshape(T, E)
    out["init_state"] = out["init_state"].reshape(E, -1).view(L.STATE_DTYPE).reshape(E)
    return out


def run(total_steps=1_000_000, seed0=500001, workers=None, episodes_per_chunk=512, one_launch=False, log=print):
    """Record >= total_steps reference steps over the four configs and replay them on CUDA.  Returns a dict
    {config: steps}; raises AssertionError on the first mismatch."""
    import torch
    import balatro_gym_b200
    from test_oracle_golden import replay
    from test_gpu_parity import CudaStepper
    assert balatro_gym_b200.load().bgym_set_option(1, (1 << 40) if one_launch else 0) == 0
    workers = workers or os.cpu_count() or 1
    t0 = time.time()
    done = {c: 0 for c in CONFIGS}
    raised = {c: 0 for c in CONFIGS}
    per_cfg = total_steps / len(CONFIGS)
    next_seed = {c: seed0 + 1_000_000 * k for k, c in enumerate(CONFIGS)}
    ctx = mp.get_context("forkserver")   # the parent is multi-threaded and holds a CUDA context: fork from a clean server
    with tempfile.TemporaryDirectory() as tmp, ctx.Pool(workers) as pool:
        while any(done[c] < per_cfg for c in CONFIGS):
            jobs = []
            for c in CONFIGS:
                if done[c] >= per_cfg:
                    continue
                # as many chunks as the remaining steps need (mean episode length ~60 steps; c4x plans are shorter)
                want = int(np.ceil((per_cfg - done[c]) / (episodes_per_chunk * (30 if c == "c4x" else 60))))
                for _ in range(max(1, min(want, 2 * workers))):
                    jobs.append((c, episodes_per_chunk, next_seed[c], tmp))
                    next_seed[c] += episodes_per_chunk + 7919      # c4x retries move seeds forward by 7919 per attempt
            for path, cfg, n_ref, n_exc in pool.imap_unordered(_record_chunk, jobs):
                tr = _load(path)
                os.remove(path)
                n = replay(tr, CudaStepper(torch, tr["action"].shape[1]))
                assert n == n_ref
                done[cfg] += n
                raised[cfg] += n_exc
    total = sum(done.values())
    for c in CONFIGS:
        log(f"config {c}: {done[c]} reference steps replayed on CUDA, {raised[c]} steps where the reference raised "
            f"(SafeBalatroEnv convention checked), 0 mismatches")
    log(f"step path: {'one-launch small-slab kernel' if one_launch else 'multi-pass (main + list kernels)'}")
    log(f"total {total} steps, 0 mismatches, wall {time.time() - t0:.0f} s, {workers} recording processes (reference = "
        f"byte-compiled copy of the unmodified sources, run live on this box; comparison = tests/test_oracle_golden.py::replay)")
    return done


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1_000_000)
    ap.add_argument("--seed0", type=int, default=500001)
    ap.add_argument("--workers", type=int, default=0)
    ap.add_argument("--one-launch", action="store_true", help="exercise the small-slab kernel instead of the multi-pass step")
    args = ap.parse_args()
    lines = []

    def log(s):
        print(s)
        lines.append(s)
    run(args.steps, args.seed0, args.workers or None, one_launch=args.one_launch, log=log)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "lockstep_cuda.txt"), "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
