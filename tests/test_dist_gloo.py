"""CPU, world_size 2 over gloo: the N > 1 host logic - env sharding and the advantage-statistics all-reduce give
every shard the global mean/std (sharded result == single-process result on the concatenated envs)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from massive_marl_benchmark_b200 import dist as mdist
    r, w, lr = mdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    gen = torch.Generator().manual_seed(0)
    T, N = 5, 37
    adv = torch.randn(T, N, generator=gen, dtype=torch.float64)
    lo, hi = mdist.shard_range(N, rank, world)
    mine = adv[:, lo:hi]
    stats = torch.tensor([mine.numel(), mine.sum(), (mine * mine).sum()], dtype=torch.float64)
    mdist.all_reduce_stats(stats, True)
    mean = stats[1] / stats[0]
    std = ((stats[2] - stats[1] * mean) / (stats[0] - 1)).sqrt()
    assert abs(float(mean) - float(adv.mean())) < 1e-12 and abs(float(std) - float(adv.std())) < 1e-12
    # gradient averaging helper
    p = torch.nn.Parameter(torch.zeros(3, 4))
    p.grad = torch.full((3, 4), float(rank + 1))
    q = torch.nn.Parameter(torch.zeros(5))
    q.grad = torch.full((5,), float(10 * (rank + 1)))
    mdist.all_reduce_grads([p, q], bucket_bytes=16)
    assert torch.allclose(p.grad, torch.full((3, 4), 1.5)) and torch.allclose(q.grad, torch.full((5,), 15.0))
    assert mdist.max_over_ranks(float(rank), "cpu") == 1.0 and mdist.sum_over_ranks(1.0, "cpu") == 2.0
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("1")


def test_stats_allreduce_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
