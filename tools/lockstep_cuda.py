#!/usr/bin/env python
"""Reference-vs-CUDA differential run on the GPU box, with fresh seeds (not the committed golden traces):
the unmodified reference (byte-compiled copy under oracle/_ref) plays episodes with every RNG draw tapped,
then the CUDA env replays the same actions and draws through the C-ABI and every step is compared —
state record, observation record, reward, termination, score breakdown, error flag.

    python tools/lockstep_cuda.py --episodes 600 --seed0 500001
Writes gpurun_out/lockstep_cuda.txt.
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--episodes", type=int, default=600, help="episodes per configuration")
    ap.add_argument("--seed0", type=int, default=500001)
    ap.add_argument("--max-steps", type=int, default=600)
    ap.add_argument("--one-launch", action="store_true", help="exercise the small-slab kernel instead of the multi-pass step")
    args = ap.parse_args()
    import torch
    import balatro_gym_b200
    assert balatro_gym_b200.load().bgym_set_option(1, (1 << 40) if args.one_launch else 0) == 0
    from make_golden import record
    from test_oracle_golden import replay
    from test_gpu_parity import CudaStepper

    lines = []
    total = 0
    t0 = time.time()
    for k, cfg in enumerate(("c1", "c3", "c4")):
        tr = record(cfg, args.episodes, args.seed0 + 100000 * k, args.max_steps)
        n_ref = int(tr["length"].sum())
        n = replay(tr, CudaStepper(torch, tr["action"].shape[1]))
        assert n == n_ref
        total += n
        lines.append(f"config {cfg}: {args.episodes} episodes (seeds {args.seed0 + 100000 * k}..), {n} reference steps replayed on CUDA, "
                     f"{int(tr['exc'].sum())} steps where the reference raised (SafeBalatroEnv convention checked), 0 mismatches")
    lines.append(f"step path: {'one-launch small-slab kernel' if args.one_launch else 'multi-pass (main + gather kernels)'}")
    lines.append(f"total {total} steps, 0 mismatches, wall {time.time() - t0:.0f} s (reference = byte-compiled copy of the unmodified "
                 f"sources, run live on this box; comparison = tests/test_oracle_golden.py::replay)")
    msg = "\n".join(lines)
    print(msg)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "lockstep_cuda.txt"), "w").write(msg + "\n")


if __name__ == "__main__":
    main()
