"""GPU: the drop-in `PPO.update` (massive_marl_benchmark_b200.ppo_update: device shuffle -> fused gather -> torch MLPs ->
fused loss kernel -> backward / clip / step) against oracle.ppo_loss_oracle.ppo_update_oracle, the CPU restatement that
tests/test_oracle_vs_reference.py pins bit for bit against the reference's own unmodified `PPO.update`.

Plain SGD for the parameter comparison: the first Adam steps are lr * g / (|g| + 1e-8), which turns rounding-level
differences of near-zero gradients into full-size steps and says nothing about the path under test; Adam (the reference's
optimiser) is exercised by the second test on the loss trajectory.  Tolerance 2e-4 of each tensor's update: fp32 MLP
forward + backward on the GPU (cuBLAS) vs the CPU differ at the 1e-6 level per GEMM and the update accumulates six
minibatch steps."""
import copy
import types

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


class _ActorCritic(nn.Module):
    """The attribute surface of the reference's ActorCritic that `PPO.update` touches (module.py:25-55: .actor, .critic,
    .log_std), small."""

    def __init__(self, obs_dim, A, hidden=(64, 32), init_std=0.8):
        super().__init__()

        def mlp(out):
            layers, d = [], obs_dim
            for h in hidden:
                layers += [nn.Linear(d, h), nn.ELU()]
                d = h
            return nn.Sequential(*layers, nn.Linear(d, out))

        self.actor, self.critic = mlp(A), mlp(1)
        self.log_std = nn.Parameter(torch.log(torch.tensor(init_std)) * torch.ones(A))


def _fill(st, ac, T, N, obs_dim, A, seed):
    g = torch.Generator().manual_seed(seed)
    obs = torch.randn(T, N, obs_dim, generator=g)
    with torch.no_grad():
        mu = ac.actor(obs.view(-1, obs_dim)).view(T, N, A) + 0.02 * torch.randn(T, N, A, generator=g)
        val = ac.critic(obs.view(-1, obs_dim)).view(T, N, 1)
        std = ac.log_std.exp() * ac.log_std.exp()
        act = mu + std * torch.randn(T, N, A, generator=g)
        logp = (-0.5 * (((act - mu) / std) ** 2).sum(-1) - std.log().sum() - 0.5 * A * 1.8378770664093453).view(T, N, 1)
    values = val + 0.1 * torch.randn(T, N, 1, generator=g)
    returns = val + 0.5 * torch.randn(T, N, 1, generator=g)
    adv = returns - values
    adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    fields = dict(observations=obs, actions=act, actions_log_prob=logp, values=values, returns=returns, advantages=adv,
                  mu=mu, sigma=ac.log_std.detach().repeat(T, N, 1))
    for k, v in fields.items():
        getattr(st, k).copy_(v)
    return fields


def _ppo(st, ac, opt, **kw):
    cfg = dict(storage=st, actor_critic=ac, optimizer=opt, num_mini_batches=3, num_learning_epochs=2, clip_param=0.2,
               value_loss_coef=2.0, entropy_coef=0.001, use_clipped_value_loss=True, desired_kl=0.016, schedule="adaptive",
               step_size=2e-3, max_grad_norm=1.0, asymmetric=False)
    cfg.update(kw)
    return types.SimpleNamespace(**cfg)


@pytest.mark.parametrize("sampler", ["sequential", "random"])
def test_ppo_update_matches_the_reference_pinned_oracle(cuda_device, sampler):
    from massive_marl_benchmark_b200.ppo_update import ppo_update
    from massive_marl_benchmark_b200.storage import RolloutStorage
    from oracle.ppo_loss_oracle import ppo_update_oracle
    dev = cuda_device
    T, N, obs_dim, A = 8, 96, 24, 8
    torch.manual_seed(11)
    ac_cpu = _ActorCritic(obs_dim, A)
    ac_gpu = copy.deepcopy(ac_cpu).to(dev)
    initial = copy.deepcopy(ac_cpu.state_dict())

    orders = [torch.arange(T * N)] * 2
    if sampler == "random":                        # host-drawn orders, replayed on both sides (the device-side permutation
        gen = torch.Generator().manual_seed(5)     # kernel has its own tests)
        orders = [torch.randperm(T * N, generator=gen) for _ in range(2)]

    class _Replayed(RolloutStorage):
        def _epoch_order(self):
            return self._orders.pop(0).to(dev)

    st_gpu = _Replayed(N, T, (obs_dim,), (0,), (A,), dev, sampler)
    st_gpu._orders = list(orders)
    st_cpu = types.SimpleNamespace(num_envs=N, num_transitions_per_env=T, states=torch.zeros(T, N, 0))
    fields = _fill(st_gpu, ac_cpu, T, N, obs_dim, A, seed=21)
    for k, v in fields.items():
        setattr(st_cpu, k, v.clone())

    gpu = _ppo(st_gpu, ac_gpu, torch.optim.SGD(ac_gpu.parameters(), lr=1e-2))
    out_gpu = ppo_update(gpu)
    cpu = _ppo(st_cpu, ac_cpu, torch.optim.SGD(ac_cpu.parameters(), lr=1e-2))
    out_cpu = ppo_update_oracle(cpu, orders)

    assert gpu.step_size == cpu.step_size != 2e-3             # the same adaptive-KL decisions on every minibatch
    assert out_gpu == pytest.approx(out_cpu, rel=1e-4, abs=2e-6)     # (the surrogate mean is a cancelling sum of O(1) terms)
    moved = 0
    for k, w0 in initial.items():
        d_cpu = ac_cpu.state_dict()[k] - w0
        d_gpu = ac_gpu.state_dict()[k].cpu() - w0
        scale = float(d_cpu.abs().max())
        ulp = 1.1920929e-07 * float(w0.abs().max())             # the parameter itself is only representable to this
        assert float((d_gpu - d_cpu).abs().max()) <= 2e-4 * scale + 2 * ulp + 1e-9, k
        moved += scale > 0
    assert moved == len(initial)


def test_ppo_update_with_adam_reduces_the_loss(cuda_device):
    """The reference's optimiser (Adam, ppo.py:73) over a few updates on a fixed rollout: finite parameters and a
    decreasing value loss; works on a reference-style storage (no fused gather) as well."""
    from massive_marl_benchmark_b200.ppo_update import ppo_update
    from massive_marl_benchmark_b200.storage import RolloutStorage
    dev = cuda_device
    T, N, obs_dim, A = 8, 128, 24, 8
    torch.manual_seed(12)
    ac = _ActorCritic(obs_dim, A)
    st = RolloutStorage(N, T, (obs_dim,), (0,), (A,), dev, "random")
    _fill(st, ac, T, N, obs_dim, A, seed=22)
    ac = ac.to(dev)
    ppo = _ppo(st, ac, torch.optim.Adam(ac.parameters(), lr=3e-4), step_size=3e-4, num_mini_batches=4)
    losses = [ppo_update(ppo)[0] for _ in range(6)]
    assert all(torch.isfinite(p).all() for p in ac.parameters())
    assert losses[-1] < losses[0]

    class _RefStyle:                                # plain tensors + a list-yielding generator, like storage.py:75-87
        pass
    ref_st = _RefStyle()
    for k in RolloutStorage.FIELDS:
        setattr(ref_st, k, getattr(st, k))
    ref_st.mini_batch_generator = lambda n: [list(range(i, T * N, n)) for i in range(n)]
    ppo.storage = ref_st
    v, s = ppo_update(ppo)
    assert v == v and s == s
