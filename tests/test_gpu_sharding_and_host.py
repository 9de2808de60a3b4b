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
