"""GPU: (1) env sharding - two shards processed independently (as two ranks would) reproduce the single run on the
concatenated envs bit for bit, and summing their (count, sum, sumsq) gives the same normalised advantages;
(2) the pinned-host provider (copy-stream prefetch) yields exactly the device-resident provider's results."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cfg(N):
    return {"env": {"numEnvs": N, "env_name": "ten_ant"}, "sim": {"dt": 0.0166}, "seed": 3}


def test_env_sharding_equals_single_run(cuda_device):
    from massive_marl_benchmark_b200 import dist as mdist
    from massive_marl_benchmark_b200 import synthetic
    from massive_marl_benchmark_b200.providers import ReplayProvider
    from massive_marl_benchmark_b200.storage import RolloutStorage
    from massive_marl_benchmark_b200.tasks import TenAnt
    dev = cuda_device
    N, T, G = 96, 6, 2
    fr = synthetic.ten_ant_frames(N, T, seed=12, fall_prob=0.01)
    frd = {k: v.to(dev) for k, v in fr.items()}
    values = torch.randn(T, N, 1, device=dev); last_values = torch.randn(N, 1, device=dev)

    def run(lo, hi):
        n = hi - lo
        sub = {"root": frd["root"].view(T, N, 143)[:, lo:hi].reshape(T, 11 * n, 13).contiguous(),
               "dof": frd["dof"].view(T, N, 160)[:, lo:hi].reshape(T, 80 * n, 2).contiguous()}
        act = frd["actions"][:, lo:hi].contiguous()
        task = TenAnt(_cfg(n), provider=ReplayProvider({k: v.cpu() for k, v in sub.items()}, device=dev))
        task.clip_actions, task.clip_obs = 1.0, 5.0
        st = RolloutStorage(n, T, (388,), (0,), (80,), dev)
        st.values.copy_(values[:, lo:hi])
        task.replay(sub, act, st.obs_slots[1:], st.rewards.view(T, n), st.dones.view(T, n))
        st.compute_returns_scan(last_values[lo:hi].contiguous(), 0.96, 0.95)
        return task, st

    _, full = run(0, N)
    stats_full = full.adv_stats.clone()
    full.normalize_advantages()
    shards = [run(*mdist.shard_range(N, r, G)) for r in range(G)]
    total = sum(st.adv_stats for _, st in shards)          # what the NCCL all-reduce of the 3 doubles computes
    assert torch.allclose(total, stats_full, rtol=1e-12, atol=1e-9)
    lo = 0
    for (task, st) in shards:
        n = st.num_envs
        assert torch.equal(st.obs_slots[1:], full.obs_slots[1:, lo:lo + n])
        assert torch.equal(st.rewards, full.rewards[:, lo:lo + n]) and torch.equal(st.dones, full.dones[:, lo:lo + n])
        assert torch.equal(st.returns, full.returns[:, lo:lo + n])
        st.adv_stats.copy_(total)
        st.normalize_advantages()
        assert torch.allclose(st.advantages, full.advantages[:, lo:lo + n], rtol=1e-6, atol=1e-6)
        lo += n


def test_host_provider_equals_device_provider(cuda_device):
    from massive_marl_benchmark_b200 import synthetic
    from massive_marl_benchmark_b200.providers import HostReplayProvider, ReplayProvider
    from massive_marl_benchmark_b200.tasks import TenAnt
    from massive_marl_benchmark_b200.vec_task import VecTaskPython
    dev = cuda_device
    N, T = 200, 9
    fr = synthetic.ten_ant_frames(N, T, seed=8, fall_prob=0.01)
    frames = {"root": fr["root"], "dof": fr["dof"]}
    a = fr["actions"].to(dev)
    env_d = VecTaskPython(TenAnt(_cfg(N), provider=ReplayProvider(frames, device=dev)), dev)
    env_h = VecTaskPython(TenAnt(_cfg(N), provider=HostReplayProvider(frames, dev)), dev)
    for t in range(2 * T):          # wraps around the replay twice (prefetch + staging reuse)
        od, rd, dd, _ = env_d.step(a[t % T])
        oh, rh, dh, _ = env_h.step(a[t % T])
        assert torch.equal(od, oh) and torch.equal(rd, rh) and torch.equal(dd, dh), t
        assert torch.equal(env_d.task.obs_buf, env_h.task.obs_buf)


# ---------------------------------------------------------------------------------------------------------
# peer-memory statistics exchange (mmb_xchg): the protocol with several endpoints inside one process
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2, 5])
def test_stats_exchange_equals_unsharded_normalisation(cuda_device, world):
    """`world` env shards, each with its own storage + exchange endpoint (mailboxes in local memory), must produce
    the advantages of ONE unsharded storage: same returns bit-for-bit, normalised advantages equal to fp32 rounding
    of the fp64 moments (the summation order over shards differs from the single accumulator's)."""
    from massive_marl_benchmark_b200.dist import StatsExchange, shard_range
    from massive_marl_benchmark_b200.storage import RolloutStorage
    dev = cuda_device
    T, N = 8, 1000
    gen = torch.Generator().manual_seed(5)
    rewards = torch.randn(T, N, 1, generator=gen).to(dev)
    values = torch.randn(T, N, 1, generator=gen).to(dev)
    dones = (torch.rand(T, N, 1, generator=gen) < 0.05).to(torch.uint8).to(dev)
    last = torch.randn(N, 1, generator=gen).to(dev)

    def fill(st, lo, hi):
        st.rewards.copy_(rewards[:, lo:hi]); st.values.copy_(values[:, lo:hi]); st.dones.copy_(dones[:, lo:hi])

    whole = RolloutStorage(N, T, (4,), (0,), (2,), dev)
    fill(whole, 0, N)
    whole.compute_returns(last, 0.99, 0.95)

    ends = StatsExchange.local(world, slots=4, device=dev)
    shards = []
    for r in range(world):
        lo, hi = shard_range(N, r, world)
        st = RolloutStorage(hi - lo, T, (4,), (0,), (2,), dev)
        st.stats_exchange = ends[r]
        shards.append((st, lo, hi))
    streams = [torch.cuda.Stream() for _ in range(world)]
    for rep in range(6):                      # more exchanges than slots: the ring wraps
        for st, lo, hi in shards:
            fill(st, lo, hi)
            st.compute_returns_scan(last[lo:hi], 0.99, 0.95)
        torch.cuda.synchronize()
        for (st, lo, hi), stream in zip(shards, streams):     # one stream per emulated rank: each normalise launch
            with torch.cuda.stream(stream):                   # waits for the others' publications
                st.normalize_advantages()
        torch.cuda.synchronize()
        for st, lo, hi in shards:
            assert torch.equal(st.returns, whole.returns[:, lo:hi])
            assert torch.allclose(st.advantages, whole.advantages[:, lo:hi], rtol=2e-6, atol=2e-6)
        # every rank normalised with bit-identical moments: shard 0's mean/std reproduce on all shards
    assert all(e.errors == 0 for e in ends)
    assert all(int(e.state[1]) == 6 for e in ends)
    assert all(float(st._adv_stats4.abs().sum()) == 0.0 for st, _, _ in shards)   # accumulator cleared by the exchange


@pytest.mark.gpu
def test_stats_exchange_streams_and_graph_replay(cuda_device):
    """Normalise kernels enqueued BEFORE the peer's scan (on other streams) wait on the mailbox flag; the whole
    pattern replays from a CUDA graph (device-side sequence counters, no host arguments change)."""
    from massive_marl_benchmark_b200.dist import StatsExchange
    from massive_marl_benchmark_b200.storage import RolloutStorage
    dev = cuda_device
    T, N = 4, 512
    ends = StatsExchange.local(2, slots=4, device=dev)
    sts = [RolloutStorage(N, T, (4,), (0,), (2,), dev) for _ in range(2)]
    last = torch.randn(N, 1, device=dev)
    ref = RolloutStorage(2 * N, T, (4,), (0,), (2,), dev)
    for r, st in enumerate(sts):
        st.stats_exchange = ends[r]
        st.rewards.normal_(); st.values.normal_()
        ref.rewards[:, r * N:(r + 1) * N] = st.rewards; ref.values[:, r * N:(r + 1) * N] = st.values
    ref.compute_returns(torch.cat([last, last]), 0.99, 0.95)
    s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda._sleep(1)     # load the delay kernel now: a first launch (lazy module load) would block behind the spinning kernel
    torch.cuda.synchronize()
    # rank 0 scans and immediately normalises on s0 (must wait ~ for rank 1), rank 1 follows later on s1
    with torch.cuda.stream(s0):
        sts[0].compute_returns(last, 0.99, 0.95)
    with torch.cuda.stream(s1):
        torch.cuda._sleep(20_000_000)          # ~10 ms: rank 0's normalise is spinning by now
        sts[1].compute_returns(last, 0.99, 0.95)
    torch.cuda.synchronize()
    for r, st in enumerate(sts):
        assert torch.allclose(st.advantages, ref.advantages[:, r * N:(r + 1) * N], rtol=2e-6, atol=2e-6)
    assert ends[0].errors == 0 and ends[1].errors == 0

    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        main = torch.cuda.current_stream()
        for st in sts:
            st.compute_returns_scan(last, 0.99, 0.95)
        s1.wait_stream(main)
        sts[0].normalize_advantages()           # the two ranks' exchange kernels on parallel graph branches
        with torch.cuda.stream(s1):
            sts[1].normalize_advantages()
        main.wait_stream(s1)
    for _ in range(9):                          # wraps the 4-slot ring twice
        g.replay()
    torch.cuda.synchronize()
    for r, st in enumerate(sts):
        assert torch.allclose(st.advantages, ref.advantages[:, r * N:(r + 1) * N], rtol=2e-6, atol=2e-6)
    assert ends[0].errors == 0 and ends[1].errors == 0


@pytest.mark.gpu
def test_stats_exchange_ipc_endpoint_single_rank(cuda_device):
    """The IPC-allocated endpoint (world 1: alloc + handle export, no peers) behaves like the plain path."""
    from massive_marl_benchmark_b200.dist import StatsExchange
    from massive_marl_benchmark_b200.storage import RolloutStorage
    dev = cuda_device
    x = StatsExchange(slots=2, device=dev)
    a, b = (RolloutStorage(300, 5, (4,), (0,), (2,), dev) for _ in range(2))
    a.rewards.normal_(); a.values.normal_()
    b.rewards.copy_(a.rewards); b.values.copy_(a.values)
    b.stats_exchange = x
    last = torch.randn(300, 1, device=dev)
    for _ in range(3):
        a.compute_returns(last, 0.99, 0.95)
        b.compute_returns(last, 0.99, 0.95)
    torch.cuda.synchronize()
    assert torch.equal(a.returns, b.returns) and torch.equal(a.advantages, b.advantages)
    assert x.errors == 0
    x.close()


@pytest.mark.gpu
def test_stats_exchange_timeout_is_loud(cuda_device):
    """A peer that never publishes: after the (configurable) time-out the exchange does NOT normalise with partial moments -
    the shard's advantages become NaN, the endpoint counts an error and `RolloutStorage.check_exchange` (called from
    `mini_batch_generator`, once per update) raises."""
    from massive_marl_benchmark_b200 import _lib as L
    from massive_marl_benchmark_b200.dist import StatsExchange
    from massive_marl_benchmark_b200.storage import RolloutStorage
    dev = cuda_device
    ends = StatsExchange.local(2, slots=4, device=dev, timeout_ms=30)
    st = RolloutStorage(256, 4, (4,), (0,), (2,), dev)
    st.stats_exchange = ends[0]
    st.rewards.normal_(); st.values.normal_()
    st.compute_returns(torch.randn(256, 1, device=dev), 0.99, 0.95)       # rank 1 stays silent
    torch.cuda.synchronize()
    assert torch.isnan(st.advantages).all(), "a failed exchange must not leave plausible numbers behind"
    assert ends[0].errors == 1
    with pytest.raises(L.MmbError, match="exchange failed"):
        st.mini_batch_generator(4)
    assert torch.isfinite(st.returns).all()
