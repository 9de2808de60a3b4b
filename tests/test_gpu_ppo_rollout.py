"""`ppo_rollout.GraphedPPORollout` (the rollout phase of ppo.py:127-157 as one CUDA-graph replay) against the same loop
written out eagerly on the replacement classes - which are themselves pinned to the reference (test_gpu_ten_ant.py,
test_gpu_storage.py, test_gpu_mlp.py, test_gpu_dropin.py): every stored plane of every rollout bit-equal, the Philox
streams of the policy's sampling and of the reset noise included (device-resident counters vs the host's), across several
phases of the frame ring, with an in-place parameter update between two rollouts, for an even and an odd horizon.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

_RESET_STATE = ("reset_count", "env_ids", "ant_box_indices", "ant_indices", "dof_reset_staging")
_PLANES = ("observations", "actions", "rewards", "dones", "values", "actions_log_prob", "mu", "sigma", "returns", "advantages")


def _actor_critic(dev):
    def net(out_dim):
        dims, mods = [388, 1024, 1024, 512, out_dim], []
        for i in range(4):
            mods.append(torch.nn.Linear(dims[i], dims[i + 1]))
            if i < 3:
                mods.append(torch.nn.ELU())
        return torch.nn.Sequential(*mods)

    class AC(torch.nn.Module):          # module.py:25-55
        def __init__(self):
            super().__init__()
            self.asymmetric = False
            self.actor, self.critic = net(80), net(1)
            self.log_std = torch.nn.Parameter(torch.log(torch.tensor(0.8)) * torch.ones(80))

    torch.manual_seed(11)
    return AC().to(dev)


def _setup(N, F, T, dev, ac):
    from massive_marl_benchmark_b200 import synthetic
    from massive_marl_benchmark_b200.mlp import PPOActorCriticForward
    from massive_marl_benchmark_b200.providers import ReplayProvider
    from massive_marl_benchmark_b200.storage import RolloutStorage
    from massive_marl_benchmark_b200.tasks import TenAnt
    from massive_marl_benchmark_b200.vec_task import VecTaskPython
    fr = synthetic.ten_ant_frames(N, F, seed=3)
    cfg = {"env": {"numEnvs": N, "env_name": "ten_ant", "episodeLength": 7}, "sim": {"dt": 0.0166}, "seed": 5}
    task = TenAnt(cfg, None, None, "cuda", 0, True, False, provider=ReplayProvider({"root": fr["root"], "dof": fr["dof"]}, device=dev))
    env = VecTaskPython(task, dev)
    pol = PPOActorCriticForward(ac, dev)
    pol.seed = 9
    st = RolloutStorage(N, T, (388,), (0,), (80,), dev)
    return env, pol, st


def _perturb(ac, k):
    """an in-place parameter change, as `optimizer.step()` makes it (bumps the version counters)"""
    with torch.no_grad():
        for i, p in enumerate(ac.parameters()):
            p.mul_(1.0 + 0.01 * ((i + k) % 3))
        ac.log_std.add_(0.05)


def _snapshot(st):
    return {k: getattr(st, k).clone() for k in _PLANES}


@pytest.mark.parametrize("T,F", [(4, 8), (3, 6)])
def test_graphed_rollout_equals_the_eager_loop(cuda_device, T, F):
    from massive_marl_benchmark_b200.episodes import EpisodeTracker
    from massive_marl_benchmark_b200.ppo_rollout import GraphedPPORollout
    dev, N, R = cuda_device, 300, 7
    ac = _actor_critic(dev)
    sd0 = {k: v.clone() for k, v in ac.state_dict().items()}

    # ---- the loop of ppo.py:127-157, eager ----
    env, pol, st = _setup(N, F, T, dev, ac)
    tr_e = EpisodeTracker(N, dev)
    states = torch.zeros(N, 0, device=dev)
    torch.manual_seed(1)
    current_obs = env.reset()
    eager = []
    for r in range(R):
        if r == 4:
            _perturb(ac, r)
        for _ in range(T):
            actions, logp, values, mu, sigma = pol.act(current_obs, states)
            next_obs, rews, dones, _ = env.step(actions)
            st.add_transitions(current_obs, states, actions, rews, dones, values, logp, mu, sigma)
            assert not values.is_contiguous() and torch.equal(st.values[st.step - 1], values)      # the strided critic column
            assert torch.equal(st.mu[st.step - 1], mu) and torch.equal(st.actions_log_prob[st.step - 1, :, 0], logp)
            current_obs.copy_(next_obs)
        last_values = pol.act(current_obs, states)[2]
        st.compute_returns(last_values, 0.99, 0.95)
        tr_e.update(st.rewards, st.dones)
        eager.append((_snapshot(st), current_obs.clone()))
        st.clear()
    reset_e = env.task.reset_buf.clone(), env.task.progress_buf.clone()
    env.task.reset_idx()        # the graphed rollout has launched the NEXT step's reset compaction already (TenAnt.reset_ahead)
    lists_e = {k: getattr(env.task, k).clone() for k in _RESET_STATE}

    # ---- the same, one graph replay per rollout ----
    ac.load_state_dict(sd0)
    env, pol, st = _setup(N, F, T, dev, ac)
    tr_g = EpisodeTracker(N, dev)
    ro = GraphedPPORollout(env, pol, st, 0.99, 0.95, tracker=tr_g)
    torch.manual_seed(1)
    for r in range(R):
        if r == 4:
            _perturb(ac, r)
        ro.run()
        assert st.step == T
        snap, cur = eager[r]
        for k in _PLANES:
            assert torch.equal(getattr(st, k), snap[k]), "rollout %d: %s differs" % (r, k)
        assert torch.equal(ro.current_obs, cur), "rollout %d: current_obs differs" % r
        st.clear()
    assert ro.captures == 2          # two phases of the frame ring; rollouts 3-6 were pure replays (the parameter change lands on one)
    assert torch.equal(env.task.reset_buf, reset_e[0]) and torch.equal(env.task.progress_buf, reset_e[1])
    for k in _RESET_STATE:     # reset index lists, counts and the Philox-randomised DOF rows of the step to come
        assert torch.equal(getattr(env.task, k), lists_e[k]), k
    assert int(tr_g.finished) == int(tr_e.finished) and tr_g.deques() == tr_e.deques()
    with pytest.raises(AssertionError, match="Rollout buffer overflow"):
        st.step = 1
        ro.run()


def test_graphed_rollout_one_ant(cuda_device):
    """BASELINE configs[0] (OneAnt PPO, 64 envs): the same comparison on a task without the reset-ahead split."""
    from massive_marl_benchmark_b200 import synthetic
    from massive_marl_benchmark_b200.mlp import PPOActorCriticForward
    from massive_marl_benchmark_b200.ppo_rollout import GraphedPPORollout
    from massive_marl_benchmark_b200.providers import ReplayProvider
    from massive_marl_benchmark_b200.storage import RolloutStorage
    from massive_marl_benchmark_b200.tasks import OneAnt
    from massive_marl_benchmark_b200.vec_task import VecTaskPython
    dev, N, T, F, R = cuda_device, 64, 8, 16, 6

    def net(out_dim):
        return torch.nn.Sequential(torch.nn.Linear(60, 256), torch.nn.ELU(), torch.nn.Linear(256, 256), torch.nn.ELU(),
                                   torch.nn.Linear(256, out_dim))

    class AC(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.asymmetric = False
            self.actor, self.critic = net(8), net(1)
            self.log_std = torch.nn.Parameter(torch.log(torch.tensor(0.8)) * torch.ones(8))

    torch.manual_seed(2)
    ac = AC().to(dev)

    def setup():
        fr = synthetic.one_ant_frames(N, F, seed=3)
        cfg = {"env": {"numEnvs": N, "env_name": "one_ant", "episodeLength": 9}, "sim": {"dt": 0.0166}, "seed": 5}
        task = OneAnt(cfg, None, None, "cuda", 0, True, False, provider=ReplayProvider(fr, device=dev))
        pol = PPOActorCriticForward(ac, dev)
        pol.seed = 3
        return VecTaskPython(task, dev), pol, RolloutStorage(N, T, (60,), (0,), (8,), dev)

    env, pol, st = setup()
    states = torch.zeros(N, 0, device=dev)
    torch.manual_seed(1)
    current_obs = env.reset()
    eager = []
    for r in range(R):
        for _ in range(T):
            actions, logp, values, mu, sigma = pol.act(current_obs, states)
            next_obs, rews, dones, _ = env.step(actions)
            st.add_transitions(current_obs, states, actions, rews, dones, values, logp, mu, sigma)
            current_obs.copy_(next_obs)
        st.compute_returns(pol.act(current_obs, states)[2], 0.99, 0.95)
        eager.append(_snapshot(st))
        st.clear()
    env, pol, st = setup()
    ro = GraphedPPORollout(env, pol, st, 0.99, 0.95)
    torch.manual_seed(1)
    for r in range(R):
        ro.run()
        for k in _PLANES:
            assert torch.equal(getattr(st, k), eager[r][k]), "rollout %d: %s differs" % (r, k)
        st.clear()
    assert ro.captures == 2 and not ro._ahead


def test_gaussian_act_device_step_counter(cuda_device):
    """`step_counter`: the launch takes the Philox step from device memory and advances it - the draws equal those of the
    host-side counter started at the same value, also when the calls are replayed from a CUDA graph."""
    from massive_marl_benchmark_b200.mlp import gaussian_act
    dev = cuda_device
    gen = torch.Generator().manual_seed(0)
    M, A = 1000, 80
    mean = torch.randn(M, A, generator=gen).to(dev)
    std = (0.3 + torch.rand(A, generator=gen)).to(dev)
    host = [gaussian_act(mean, std, seed=4, step=s)[0] for s in range(5, 11)]
    from massive_marl_benchmark_b200 import _lib as L
    ctr = torch.zeros(L.ACT_COUNTER_WORDS, dtype=torch.int64, device=dev)
    ctr[0] = 5
    for s in range(2):
        assert torch.equal(gaussian_act(mean, std, seed=4, step_counter=ctr)[0], host[s])
    assert ctr.tolist() == [7] + [0] * (L.ACT_COUNTER_WORDS - 1)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = gaussian_act(mean, std, seed=4, step_counter=ctr)[0]
    for s in range(2, 6):
        g.replay()
        assert torch.equal(out, host[s]), s
    assert ctr.tolist() == [11] + [0] * (L.ACT_COUNTER_WORDS - 1)
    small = gaussian_act(mean[:3], std, seed=4, step_counter=ctr)[0]          # one block: fewer blocks than ticket lanes
    assert torch.equal(small, gaussian_act(mean[:3], std, seed=4, step=11)[0]) and ctr.tolist() == [12] + [0] * (L.ACT_COUNTER_WORDS - 1)


@pytest.mark.parametrize("M,A", [(4099, 80), (40961, 8), (517, 6), (300, 64), (1000, 7), (5, 80), (70, 200)])
def test_gaussian_act_shapes_and_paths(cuda_device, M, A):
    """Every row-grouping of the pair path (rows per warp 1 ... 16, ragged last group, widths below / above / multiple of a
    warp's 32 pairs) and the element-per-lane path (odd width) against torch with supplied draws; and the Philox stream does
    not depend on the path: an odd row stride of the means forces the element path, the draws stay the same."""
    from massive_marl_benchmark_b200.mlp import gaussian_act
    from torch.distributions import Normal
    dev = cuda_device
    gen = torch.Generator().manual_seed(M + A)
    mean = torch.randn(M, A, generator=gen).to(dev)
    std = (0.3 + torch.rand(A, generator=gen)).to(dev)
    z = torch.randn(M, A, generator=gen).to(dev)
    src = torch.randn(A, generator=gen).to(dev)
    act, lp, sig = gaussian_act(mean, std, noise=z, sigma_src=src)
    assert torch.allclose(act, mean + z * std, rtol=1e-6, atol=1e-6) and torch.equal(sig, src.repeat(M, 1))
    ref = Normal(mean, std).log_prob(act)
    assert torch.allclose(lp, ref.sum(1), rtol=1e-4, atol=2e-3)
    act2, lp2 = gaussian_act(mean, std, noise=z, per_dim=True)
    assert torch.equal(act2, act) and torch.allclose(lp2, ref, rtol=1e-4, atol=1e-4)
    groups = 3
    rows = (M + groups - 1) // groups
    std_g = (0.3 + torch.rand(groups, A, generator=gen)).to(dev)
    act3, lp3 = gaussian_act(mean, std_g, noise=z, std_group_rows=rows, per_dim=True)
    sd_rows = std_g.repeat_interleave(rows, dim=0)[:M]
    assert torch.allclose(act3, mean + z * sd_rows, rtol=1e-6, atol=1e-6)
    assert torch.allclose(lp3, Normal(mean, sd_rows).log_prob(act3), rtol=1e-4, atol=1e-4)
    a_pair, l_pair = gaussian_act(mean, std, seed=3, step=9)
    wide = torch.zeros(M, A + 1, device=dev)
    wide[:, :A] = mean
    a_elem, l_elem = gaussian_act(wide[:, :A], std, seed=3, step=9)          # row stride A + 1
    if A % 2 == 0:
        assert wide[:, :A].stride(0) % 2 == 1
    assert torch.equal(a_pair, a_elem) and torch.allclose(l_pair, l_elem, rtol=1e-5, atol=1e-5)
