#!/usr/bin/env python
"""PPO on the device vector env — the counterpart of the reference's `hpc_train.py` (args :179-195,
PPO hyper-parameters :73-90) with envs, rollout buffer, sampling and GAE on the GPU.

    python tools/train_ppo.py --n-envs 65536 --timesteps 50000000
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_ppo.py --n-envs 524288

One process per GPU; every rank owns `--n-envs` envs (global env index = rank * n_envs + i, so a
run's episodes do not depend on how many GPUs share it) and the gradient is averaged over ranks.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-envs", type=int, default=1 << 16, help="envs per GPU")
    ap.add_argument("--timesteps", type=int, default=20_000_000, help="total env-steps over all ranks")
    ap.add_argument("--n-steps", type=int, default=128)
    ap.add_argument("--learning-rate", type=float, default=3e-4)
    ap.add_argument("--batch-size", type=int, default=1 << 16, help="minibatch of the PPO update")
    ap.add_argument("--n-epochs", type=int, default=4)
    ap.add_argument("--gamma", type=float, default=0.99)
    ap.add_argument("--gae-lambda", type=float, default=0.95)
    ap.add_argument("--ent-coef", type=float, default=0.01)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--save", default=None, help="path of the final policy checkpoint (rank 0)")
    args = ap.parse_args()

    import torch
    from balatro_gym_b200 import BalatroVecEnv, dist as bdist
    from balatro_gym_b200.rollout import RolloutCollector, make_policy, ppo_update

    rank, local_rank, ws = bdist.init_process_group("nccl")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    vec = BalatroVecEnv(args.n_envs, device=dev, seed=args.seed, env_offset=rank * args.n_envs)
    vec.reset()
    policy = make_policy(device=dev, seed=args.seed)
    opt = torch.optim.Adam(policy.parameters(), lr=args.learning_rate, eps=1e-5)
    roll = RolloutCollector(vec, policy, n_steps=args.n_steps, gamma=args.gamma, gae_lambda=args.gae_lambda, seed=args.seed)
    gen = torch.Generator(device=dev).manual_seed(args.seed + rank)
    per_iter = ws * args.n_envs * args.n_steps
    done_steps, it, t0 = 0, 0, time.time()
    while done_steps < args.timesteps:
        tc = time.time()
        roll.collect()
        torch.cuda.synchronize(dev)
        t_collect = time.time() - tc
        out = ppo_update(policy, opt, roll, n_epochs=args.n_epochs, minibatch=args.batch_size, ent_coef=args.ent_coef, generator=gen)
        torch.cuda.synchronize(dev)
        steps, episodes, mean_r = roll.stats()
        s = torch.tensor([episodes, mean_r * steps, steps], dtype=torch.float64, device=dev)
        bdist.allreduce_stats(s)
        done_steps += per_iter; it += 1
        if rank == 0:
            print(json.dumps({"iter": it, "env_steps": done_steps, "episodes": int(s[0]), "mean_step_reward": float(s[1] / s[2]),
                              "collect_steps_per_s": per_iter / t_collect, "wall_s": round(time.time() - t0, 2), **out}), flush=True)
    if rank == 0 and args.save:
        torch.save(policy.state_dict(), args.save)
    bdist.barrier()
    if ws > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
