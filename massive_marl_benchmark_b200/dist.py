"""Env-sharded data parallelism (one process per GPU, torch.distributed over NCCL/NVLink).

The reference has no multi-GPU code.  The hot path shards naturally: every op of
the task pipeline is row-wise over environments, so GPU g of G owns envs [g*N/G, (g+1)*N/G) with its own
frame provider, buffers and rollout storage, and NO data-path collective is needed for obs / reward /
reset / GAE.  Only two reductions cross environments:

  1. advantage normalisation: (count, sum, sumsq) - three fp64 per storage (per agent for MARL) -
     `all_reduce_stats` between mmb_gae_* and mmb_adv_normalize;
  2. the PPO/MAPPO gradient all-reduce after backward - `all_reduce_grads` (sum / world).

Both are plain NCCL all-reduces: neither follows a compute kernel closely enough to fuse (the first is
24 bytes, latency-bound; the second belongs to autograd, out of scope for the kernels).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun); returns
    (rank, world, local_rank).  World size 1 without the env vars -> no process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank


def shard_range(num_envs: int, rank: int, world: int):
    """Contiguous env shard of `rank`: [lo, hi).  Remainder envs go to the lowest ranks."""
    base, rem = divmod(num_envs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_stats(stats: torch.Tensor, group=None):
    """Sum the (count, sum, sumsq) fp64 triples over the env shards, in place."""
    if dist.is_initialized() and dist.get_world_size(group if group is not True else None) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=None if group is True else group)
    return stats


def all_reduce_grads(parameters, group=None, bucket_bytes=64 << 20):
    """Average gradients over the shards: flattened into a few large buckets (sized for launch latency and
    overlap, not link count - NVSwitch gives every peer full bandwidth)."""
    if not dist.is_initialized():
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    grads = [p.grad for p in parameters if p.grad is not None]
    bucket, size = [], 0
    def flush():
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
            bucket, size = [], 0
    flush()


def max_over_ranks(value: float, device) -> float:
    """Device-timed durations are reported as the max over ranks."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
