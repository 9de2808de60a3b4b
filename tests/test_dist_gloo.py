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


def _update_worker(rank, world, port, out_dir):
    """Env-sharded PPO update: each rank updates on its own envs with averaged gradients / KL; the parameters must equal
    the single-process update on the concatenated minibatches."""
    import copy
    import sys
    import types
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import test_gpu_ppo_update as t
    from massive_marl_benchmark_b200 import dist as mdist
    from massive_marl_benchmark_b200 import ppo_update as pu
    from massive_marl_benchmark_b200.ppo_loss import PpoLossOut
    from oracle.ppo_loss_oracle import ppo_loss_terms, ppo_update_oracle

    def stand_in(mu, log_std, value, actions, old_logp, adv, tv, ret, old_mu, old_sigma, **cfg):     # the kernel's job, on CPU
        loss, s, v, kl, logp, ent = ppo_loss_terms(mu, log_std, value, actions, old_logp, adv, tv, ret, old_mu, old_sigma, **cfg)
        return PpoLossOut(loss, s.detach(), v.detach(), kl.detach(), logp.detach(), ent.detach()[0])

    pu.ppo_loss = stand_in
    mdist.init_from_env("gloo")
    T, N, obs_dim, A, n_mb = 6, 48, 12, 8, 3
    torch.manual_seed(31)
    ac_full = t._ActorCritic(obs_dim, A, hidden=(16, 8))
    ac_mine = copy.deepcopy(ac_full)
    names = ("observations", "actions", "actions_log_prob", "values", "returns", "advantages", "mu", "sigma")
    widths = {"observations": obs_dim, "actions": A, "mu": A, "sigma": A}
    full = types.SimpleNamespace(num_envs=N, num_transitions_per_env=T, states=torch.zeros(T, N, 0))
    for k in names:
        setattr(full, k, torch.zeros(T, N, widths.get(k, 1)))
    t._fill(full, ac_full, T, N, obs_dim, A, seed=41)

    lo, hi = mdist.shard_range(N, rank, world)
    n = hi - lo
    shard = types.SimpleNamespace(num_envs=n, num_transitions_per_env=T, states=torch.zeros(T, n, 0))
    for k in names:
        setattr(shard, k, getattr(full, k)[:, lo:hi].contiguous())
    mb = (T * n) // n_mb
    shard.mini_batch_generator = lambda k: [torch.arange(T * n)[i * mb:(i + 1) * mb] for i in range(k)]
    mine = t._ppo(shard, ac_mine, torch.optim.SGD(ac_mine.parameters(), lr=1e-2), num_mini_batches=n_mb)
    mine.grad_sync, mine.scalar_sync = mdist.all_reduce_grads, mdist.mean_over_ranks
    out_mine = pu.ppo_update(mine)

    # the single-process update whose minibatch b is the union of every shard's minibatch b
    def global_ids(r):
        l, h = mdist.shard_range(N, r, world)
        local = torch.arange(T * (h - l))
        return (local // (h - l)) * N + l + local % (h - l)
    per_rank = [global_ids(r) for r in range(world)]
    order = torch.cat([torch.cat([g[b * mb:(b + 1) * mb] for g in per_rank]) for b in range(n_mb)])
    whole = t._ppo(full, ac_full, torch.optim.SGD(ac_full.parameters(), lr=1e-2), num_mini_batches=n_mb)
    out_full = ppo_update_oracle(whole, [order, order])
    assert mine.step_size == whole.step_size != 2e-3
    assert abs(out_mine[0] - out_full[0]) < 1e-5 * abs(out_full[0]) and abs(out_mine[1] - out_full[1]) < 1e-6
    for a, b in zip(ac_mine.state_dict().values(), ac_full.state_dict().values()):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, "upd%d" % rank), "w").write("1")


def test_env_sharded_ppo_update_world2(tmp_path):
    port = _free_port()
    mp.spawn(_update_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "upd0").exists() and (tmp_path / "upd1").exists()


def _overlap_worker(rank, world, port, out_dir):
    """Bucketed all-reduce overlapped with backward (hooks) and the flat-slice reducer: averaged gradients equal the mean of
    the per-rank gradients, `.grad` ends up as views of the flat buffer, repeated steps and partially used networks work."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from massive_marl_benchmark_b200 import dist as mdist
    mdist.init_from_env("gloo")
    torch.manual_seed(5)
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ELU(), torch.nn.Linear(16, 16), torch.nn.ELU(), torch.nn.Linear(16, 3))
    unused = torch.nn.Linear(4, 2)                     # never takes part in the loss: its bucket gets no gradient
    params = list(net.parameters()) + list(unused.parameters())
    import copy
    ref = copy.deepcopy(net)                           # hook-free twin for the expected values
    red = mdist.OverlappedGradAllReduce(params, bucket_bytes=256)       # three buckets
    assert len(red.buckets) >= 3
    for step in range(3):
        xs = [torch.randn(8, 6, generator=torch.Generator().manual_seed(100 * step + r)) for r in range(world)]
        want = None
        for r in range(world):                         # what the average must be: every rank's gradient, computed locally
            ref.zero_grad()
            ref(xs[r]).pow(2).mean().backward()
            g = [p.grad.clone() for p in ref.parameters()]
            want = g if want is None else [a + b for a, b in zip(want, g)]
        want = [w / world for w in want]
        for p in params:
            p.grad = None
        net(xs[rank]).pow(2).mean().backward()          # hooks fire bucket by bucket while backward runs
        red.finish()
        for p, w in zip(net.parameters(), want):
            assert torch.allclose(p.grad, w, rtol=1e-6, atol=1e-8)
            lo = red.flat.data_ptr()
            assert lo <= p.grad.data_ptr() < lo + red.flat.numel() * 4, ".grad must be a view of the flat buffer"
        # a parameter without a gradient that shares a bucket with produced ones travels as zeros (documented in finish())
        assert all(p.grad is not None and float(p.grad.abs().sum()) == 0.0 for p in unused.parameters())
    assert red.calls >= 3 * 3 and red.bytes_reduced > 0
    red.remove()
    # slices of one flat buffer, asynchronously
    flat = torch.arange(40, dtype=torch.float32) * (rank + 1)
    fr = mdist.FlatGradReducer(flat)
    fr.reduce_async(0, 16)
    fr.reduce_async(16, 40)
    fr.wait()
    assert torch.allclose(flat, torch.arange(40, dtype=torch.float32) * (sum(range(1, world + 1)) / world))
    assert fr.calls == 2 and fr.bytes_reduced == 160
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, "ovl%d" % rank), "w").write("1")


def test_overlapped_gradient_allreduce_world2(tmp_path):
    port = _free_port()
    mp.spawn(_overlap_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ovl0").exists() and (tmp_path / "ovl1").exists()
