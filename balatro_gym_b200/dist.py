"""Multi-GPU plumbing: environments shard as contiguous slabs of the global env index, one slab
per rank, with NO collective on the step path (env i never reads env j).  The only exchange is an
optional all-reduce of the per-slab episode statistics, once per rollout."""
from __future__ import annotations

import os


def world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process = 1 GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_process_group(backend: str = "nccl"):
    import torch.distributed as dist
    rank, local_rank, ws = world()
    if ws > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        dist.init_process_group(backend=backend, rank=rank, world_size=ws)
    return rank, local_rank, ws


def slab(total_envs: int, rank: int, world_size: int):
    """Contiguous slab [start, start+count) of the global env index owned by `rank`
    (sizes differ by at most one; seeds are a function of the GLOBAL index, so results do not
    depend on the number of GPUs)."""
    base, rem = divmod(total_envs, world_size)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def allreduce_stats(stats):
    """Sum the small per-slab statistics vector over ranks (NCCL on GPUs, gloo on CPU tensors)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return value


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def allreduce_mean_grads(params):
    """Average the gradients of `params` over ranks with ONE all-reduce on a flat buffer (the policy-side
    exchange of a data-parallel PPO update; the env step path itself has no collective)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return
    params = [p for p in params if p.grad is not None]
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= dist.get_world_size()
    o = 0
    for p in params:
        p.grad.copy_(flat[o:o + p.numel()].view_as(p.grad))
        o += p.numel()
