"""GPU: the drop-in `MAPPO.ppo_update` (massive_marl_benchmark_b200.mappo_update: torch trunks -> fused loss kernel -> the
reference's two backward / clip / step sequences, PopArt called as often as the reference calls it) against
oracle.mappo_loss_oracle.mappo_update_oracle, the CPU restatement that tests/test_oracle_vs_reference.py pins bit for bit
against the reference's own trainer + MAPPO_Policy + PopArt + Adam.

The modules below carry the attribute surface of the reference's Actor / Critic that the update touches (actor.base,
actor.act.action_out.{fc_mean, log_std, std_x_coef, std_y_coef}, critic.base, critic.v_out; actor_critic.py:26-41,
150-161, distributions.py:94-117).  Plain SGD for the parameter comparison (see tests/test_gpu_ppo_update.py);
tolerance 2e-4 of each tensor's update."""
import copy
import types

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _trunk(inp, hidden):
    return nn.Sequential(nn.LayerNorm(inp), nn.Linear(inp, hidden), nn.ELU(), nn.LayerNorm(hidden),
                         nn.Linear(hidden, hidden), nn.ELU(), nn.LayerNorm(hidden))


class _Head(nn.Module):
    def __init__(self, hidden, A):
        super().__init__()
        self.fc_mean = nn.Linear(hidden, A)
        self.std_x_coef, self.std_y_coef = 1.0, 0.5
        self.log_std = nn.Parameter(torch.ones(A) * self.std_x_coef)


class _Actor(nn.Module):
    def __init__(self, obs_dim, A, hidden=32):
        super().__init__()
        self.base = _trunk(obs_dim, hidden)
        self.act = nn.Module()
        self.act.action_out = _Head(hidden, A)


class _Critic(nn.Module):
    def __init__(self, share_dim, hidden=32):
        super().__init__()
        self.base = _trunk(share_dim, hidden)
        self.v_out = nn.Linear(hidden, 1)


class _PopArt:
    """The reference's PopArt call protocol (popart.py:36-60) over the oracle's restatement of its update."""

    def __init__(self, state):
        self.state = state

    def __call__(self, x):
        from oracle.mappo_loss_oracle import popart_update
        m, v = popart_update(self.state, x)
        return (x - m) / torch.sqrt(v)

    def running_mean_var(self):
        s = self.state
        m = s["running_mean"] / s["debiasing_term"].clamp(min=1e-5)
        msq = s["running_mean_sq"] / s["debiasing_term"].clamp(min=1e-5)
        return m, (msq - m ** 2).clamp(min=1e-2)


def _sample(actor, critic, B, obs_dim, share_dim, A, seed):
    g = torch.Generator().manual_seed(seed)
    obs, share = torch.randn(B, obs_dim, generator=g), torch.randn(B, share_dim, generator=g)
    with torch.no_grad():
        head = actor.act.action_out
        mean = head.fc_mean(actor.base(obs)) + 0.05 * torch.randn(B, A, generator=g)
        std = torch.sigmoid(head.log_std / head.std_x_coef) * head.std_y_coef
        actions = mean + std * torch.randn(B, A, generator=g)
        old_logp = torch.distributions.Normal(mean, std).log_prob(actions)
        vals = critic.v_out(critic.base(share))
    value_preds = vals + 0.1 * torch.randn(B, 1, generator=g)
    returns = 1.0 + 2.0 * torch.randn(B, 1, generator=g)
    active = (torch.rand(B, 1, generator=g) > 0.2).float()
    adv = torch.randn(B, 1, generator=g)
    return (share, obs, None, None, actions, value_preds, returns, None, active, old_logp, adv, None, None)


def _trainer(actor, critic, popart_state, **over):
    policy = types.SimpleNamespace(actor=actor, critic=critic,
                                   actor_optimizer=torch.optim.SGD(actor.parameters(), lr=1e-2),
                                   critic_optimizer=torch.optim.SGD(critic.parameters(), lr=1e-2))
    cfg = dict(policy=policy, clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.01, max_grad_norm=10.0, huber_delta=10.0,
               _use_popart=popart_state is not None, _use_valuenorm=False, _use_max_grad_norm=True, _use_huber_loss=True,
               _use_clipped_value_loss=True, _use_value_active_masks=False, _use_policy_active_masks=False,
               _use_recurrent_policy=False, _use_naive_recurrent=False, popart=popart_state,
               value_normalizer=_PopArt(popart_state) if popart_state is not None else None)
    cfg.update(over)
    return types.SimpleNamespace(**cfg)


@pytest.mark.parametrize("over", [dict(), dict(_use_value_active_masks=True, _use_policy_active_masks=True, huber_delta=0.5),
                                  dict(nopop=True, _use_huber_loss=False, _use_clipped_value_loss=False)])
def test_mappo_update_matches_the_reference_pinned_oracle(cuda_device, over):
    from massive_marl_benchmark_b200.mappo_update import mappo_ppo_update
    from oracle.mappo_loss_oracle import mappo_update_oracle
    dev = cuda_device
    over = dict(over)
    nopop = over.pop("nopop", False)
    obs_dim, share_dim, A, B = 46, 388, 8, 512
    torch.manual_seed(4)
    actor_c, critic_c = _Actor(obs_dim, A), _Critic(share_dim)
    actor_g, critic_g = copy.deepcopy(actor_c).to(dev), copy.deepcopy(critic_c).to(dev)
    initial = [copy.deepcopy(actor_c.state_dict()), copy.deepcopy(critic_c.state_dict())]

    def state(device):
        if nopop:
            return None
        deb = torch.tensor(1.0 - 0.99999 ** 300)
        return {"running_mean": (torch.tensor([0.9]) * deb).to(device), "running_mean_sq": (torch.tensor([4.5]) * deb).to(device),
                "debiasing_term": deb.to(device)}

    cpu = _trainer(actor_c, critic_c, state("cpu"), **over)
    gpu = _trainer(actor_g, critic_g, state(dev), **over)
    for it in range(3):
        sample = _sample(actor_c, critic_c, B, obs_dim, share_dim, A, seed=200 + it)   # (drawn around the CPU copy's policy)
        out_c = mappo_update_oracle(cpu, sample)
        out_g = mappo_ppo_update(gpu, tuple(None if t is None else t.to(dev) for t in sample))
        for a, b, name in zip(out_c, out_g, ("value_loss", "critic_grad_norm", "policy_loss", "dist_entropy", "actor_grad_norm")):
            a, b = float(a), float(b)
            assert abs(a - b) <= 2e-4 * abs(a) + 2e-6, (it, name, a, b)
        assert torch.allclose(out_g[5].cpu(), out_c[5], rtol=2e-4, atol=1e-6)          # importance weights
    for nets, w0 in (((actor_c, actor_g), initial[0]), ((critic_c, critic_g), initial[1])):
        moved = 0
        for k, w in w0.items():
            d_c = nets[0].state_dict()[k] - w
            d_g = nets[1].state_dict()[k].cpu() - w
            scale = float(d_c.abs().max())
            ulp = 1.1920929e-07 * float(w.abs().max())          # the parameter itself is only representable to this
            assert float((d_g - d_c).abs().max()) <= 2e-4 * scale + 2 * ulp + 1e-9, k
            moved += scale > 0
        assert moved == len(w0)
    if not nopop:
        for k in cpu.popart:                                                             # six PopArt updates on both sides
            assert torch.allclose(gpu.popart[k].cpu(), cpu.popart[k], rtol=1e-6, atol=0), k


def test_mappo_update_refuses_what_it_does_not_cover(cuda_device):
    from massive_marl_benchmark_b200.mappo_update import mappo_ppo_update
    dev = cuda_device
    actor, critic = _Actor(6, 4).to(dev), _Critic(9).to(dev)
    sample = tuple(None if t is None else t.to(dev) for t in _sample(copy.deepcopy(actor).cpu(), copy.deepcopy(critic).cpu(), 16, 6, 9, 4, 1))
    with pytest.raises(NotImplementedError):
        mappo_ppo_update(_trainer(actor, critic, None, _use_recurrent_policy=True), sample)
    with pytest.raises(NotImplementedError):
        mappo_ppo_update(_trainer(actor, critic, None), sample[:11] + (torch.ones(16, 4, device=dev), None))


def _compare_updates(out_c, out_g, tag):
    for a, b, name in zip(out_c, out_g, ("value_loss", "critic_grad_norm", "policy_loss", "dist_entropy", "actor_grad_norm")):
        a, b = float(a), float(b)
        assert abs(a - b) <= 2e-4 * abs(a) + 2e-6, (tag, name, a, b)
    assert torch.allclose(out_g[5].cpu(), out_c[5], rtol=2e-4, atol=1e-6), tag          # importance weights


def _compare_parameters(pairs, initial):
    for nets, w0 in zip(pairs, initial):
        moved = 0
        for k, w in w0.items():
            d_c = nets[0].state_dict()[k] - w
            d_g = nets[1].state_dict()[k].cpu() - w
            scale = float(d_c.abs().max())
            ulp = 1.1920929e-07 * float(w.abs().max())
            assert float((d_g - d_c).abs().max()) <= 2e-4 * scale + 2 * ulp + 1e-9, k
            moved += scale > 0
        assert moved == len(w0)


def test_ippo_update_matches_the_reference_pinned_oracle(cuda_device):
    """`ippo_ppo_update` with the fused loss KERNEL (ippo_trainer.py:101-170: ValueNorm.update once, both value error
    terms on the same moments, the agent's own observation as the critic input) against the oracle that
    tests/test_oracle_vs_reference.py pins on the reference's IPPO trainer."""
    from massive_marl_benchmark_b200.mappo_update import ippo_ppo_update
    from oracle.mappo_loss_oracle import mappo_update_oracle, popart_update
    dev = cuda_device

    class _ValueNorm(_PopArt):                           # valuenorm.py:39-55: update() only in training mode
        def update(self, x):
            popart_update(self.state, x)

        def __call__(self, x):
            raise AssertionError("IPPO never calls the normaliser in training mode")

    obs_dim, A, B = 46, 8, 512
    torch.manual_seed(6)
    actor_c, critic_c = _Actor(obs_dim, A), _Critic(obs_dim)
    actor_g, critic_g = copy.deepcopy(actor_c).to(dev), copy.deepcopy(critic_c).to(dev)
    initial = [copy.deepcopy(actor_c.state_dict()), copy.deepcopy(critic_c.state_dict())]

    def state(device):
        return {"running_mean": torch.zeros(1, device=device), "running_mean_sq": torch.zeros(1, device=device),
                "debiasing_term": torch.tensor(0.0, device=device)}

    cpu = _trainer(actor_c, critic_c, state("cpu"), _use_popart=False, _use_valuenorm=True)
    gpu = _trainer(actor_g, critic_g, state(dev), _use_popart=False, _use_valuenorm=True)
    gpu.value_normalizer = _ValueNorm(gpu.popart)
    for it in range(3):
        sample = _sample(actor_c, critic_c, B, obs_dim, obs_dim, A, seed=400 + it)
        out_c = mappo_update_oracle(cpu, sample, ippo=True)
        out_g = ippo_ppo_update(gpu, tuple(None if t is None else t.to(dev) for t in sample))
        _compare_updates(out_c, out_g, it)
    _compare_parameters(((actor_c, actor_g), (critic_c, critic_g)), initial)
    assert float(cpu.popart["debiasing_term"]) > 0
    for k in cpu.popart:
        assert torch.allclose(gpu.popart[k].cpu(), cpu.popart[k], rtol=1e-6, atol=0), k


@pytest.mark.parametrize("over", [dict(), dict(_use_policy_active_masks=True)])
def test_happo_update_matches_the_reference_pinned_oracle(cuda_device, over):
    """`happo_ppo_update` with the fused loss KERNEL (happo_trainer.py:93-170: the sequential-update factor inside the
    surrogate; the drop-in folds a positive factor into the advantage) against the reference-pinned oracle; a per-action
    factor (first iteration) and a per-row factor."""
    from massive_marl_benchmark_b200.mappo_update import happo_ppo_update
    from oracle.mappo_loss_oracle import mappo_update_oracle
    dev = cuda_device
    obs_dim, share_dim, A, B = 46, 388, 8, 512
    torch.manual_seed(7)
    actor_c, critic_c = _Actor(obs_dim, A), _Critic(share_dim)
    actor_g, critic_g = copy.deepcopy(actor_c).to(dev), copy.deepcopy(critic_c).to(dev)
    initial = [copy.deepcopy(actor_c.state_dict()), copy.deepcopy(critic_c.state_dict())]

    def state(device):
        deb = torch.tensor(1.0 - 0.99999 ** 300)
        return {"running_mean": (torch.tensor([0.9]) * deb).to(device), "running_mean_sq": (torch.tensor([4.5]) * deb).to(device),
                "debiasing_term": deb.to(device)}

    cpu = _trainer(actor_c, critic_c, state("cpu"), **over)
    gpu = _trainer(actor_g, critic_g, state(dev), **over)
    del gpu._use_valuenorm                               # the HAPPO trainer has no such attribute (happo_trainer.py:30-42)
    for it in range(3):
        sample = _sample(actor_c, critic_c, B, obs_dim, share_dim, A, seed=700 + it)
        g = torch.Generator().manual_seed(800 + it)
        factor = torch.exp(0.3 * torch.randn(B, 1, generator=g))
        sample = sample[:12] + (factor if it else factor.repeat(1, A) / A,)
        out_c = mappo_update_oracle(cpu, sample, happo=True)
        out_g = happo_ppo_update(gpu, tuple(None if t is None else t.to(dev) for t in sample))
        _compare_updates(out_c, out_g, (over, it))
    _compare_parameters(((actor_c, actor_g), (critic_c, critic_g)), initial)
    for k in cpu.popart:
        assert torch.allclose(gpu.popart[k].cpu(), cpu.popart[k], rtol=1e-6, atol=0), k
