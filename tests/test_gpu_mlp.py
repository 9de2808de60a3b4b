"""GPU parity of the tcgen05 MLP forward (mmb_mlp_layer / mmb_ln_cast through mlp.FusedMLP).

Two references, both plain PyTorch fp32 of the same op:
  (1) the exact fp32 forward (the reference's SGEMM path; golden outputs of the reference's own MLPBase on the shipped
      checkpoint for the MARL actor) - tolerance = bf16 operand precision: |d| <= 2e-2 * max|row| (measured ~3e-3);
  (2) the same forward with operands rounded to bf16 where the kernel rounds them (activations and weights), fp32
      accumulate - isolates the kernel's arithmetic from the quantisation: |d| <= 2e-3 * max|row| (differences come
      only from accumulation order and from bf16 re-rounding of near-tie activations).
"""
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _rowmax_err(a, b):
    """max |a-b| per row relative to that row's max |b|, but never to less than the RMS of the whole output (a 1-wide
    head has rows whose only entry can be arbitrarily close to zero)."""
    floor = b.pow(2).mean().sqrt().clamp(min=1e-6)
    return float(((a - b).abs().amax(dim=1) / torch.maximum(b.abs().amax(dim=1), floor)).max())


def _bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def _ppo_net(in_dim, hidden, out_dim, gain_last, gen):
    """module.py:25-49,57-64: Linear/ELU chain, orthogonal init (gain sqrt(2), last layer gain_last)."""
    dims = [in_dim] + hidden + [out_dim]
    mods = []
    for i in range(len(dims) - 1):
        lin = torch.nn.Linear(dims[i], dims[i + 1])
        torch.nn.init.orthogonal_(lin.weight, gain=(gain_last if i == len(dims) - 2 else 2 ** 0.5))
        lin.bias.data.uniform_(-0.1, 0.1, generator=gen)
        mods.append(lin)
        if i < len(dims) - 2:
            mods.append(torch.nn.ELU())
    return torch.nn.Sequential(*mods)


@pytest.mark.parametrize("M", [1, 100, 128, 4096])
def test_ppo_actor_critic_forward(cuda_device, M):
    from massive_marl_benchmark_b200.mlp import FusedMLP
    dev = cuda_device
    torch.backends.cuda.matmul.allow_tf32 = False
    gen = torch.Generator().manual_seed(M)
    torch.manual_seed(M)
    for out_dim, gain in ((80, 0.01), (1, 1.0)):      # actor mean head, critic value head
        net = _ppo_net(388, [1024, 1024, 512], out_dim, gain, gen).to(dev)
        x = torch.clamp(torch.randn(M, 388, generator=gen) * 2.0, -5, 5).to(dev)
        fused = FusedMLP.from_sequential(net, dev)
        y = fused(x)
        torch.cuda.synchronize()
        with torch.no_grad():
            ref = net(x)
            h = _bf(x)
            lins = [m for m in net if isinstance(m, torch.nn.Linear)]
            for i, lin in enumerate(lins):
                h = F.linear(h, _bf(lin.weight), lin.bias)
                if i < len(lins) - 1:
                    h = _bf(F.elu(h))
            emu = h
        assert y.shape == ref.shape and torch.isfinite(y).all()
        # multi-layer: intermediate activations are re-rounded to bf16, and a near-tie value can round the other way
        # under a different fp32 accumulation order (1 bf16 ulp = 0.4 %), so the emulation bound is not ulp-tight
        assert _rowmax_err(y, emu) <= 1e-2, ("vs bf16-operand emulation", M, out_dim, _rowmax_err(y, emu))
        assert _rowmax_err(y, ref) <= 3e-2, ("vs fp32", M, out_dim, _rowmax_err(y, ref))


@pytest.mark.parametrize("M,K,N", [(1, 388, 80), (300, 46, 8), (128, 512, 1), (1000, 1024, 96), (257, 64, 256), (4096, 388, 32)])
def test_single_layer_gemm_exact(cuda_device, M, K, N):
    """One Linear layer, fp32 output: no intermediate rounding, so the only difference to fp32 torch on bf16-rounded
    operands is the accumulation order -> 1e-4.  This pins descriptors, swizzle, TMEM read-back and tile edges."""
    from massive_marl_benchmark_b200.mlp import FusedMLP
    dev = cuda_device
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(M + K + N)
    lin = torch.nn.Linear(K, N).to(dev)
    x = torch.randn(M, K, device=dev)
    y = FusedMLP.from_sequential(torch.nn.Sequential(lin), dev)(x)
    with torch.no_grad():
        emu = F.linear(_bf(x), _bf(lin.weight), lin.bias)
    torch.cuda.synchronize()
    assert _rowmax_err(y, emu) <= 1e-4, _rowmax_err(y, emu)


def test_marl_actor_shipped_checkpoint(cuda_device):
    """Real weights: the shipped TenAnt MAPPO actor of agent 0; expected output = the reference's own MLPBase (fp32)."""
    from massive_marl_benchmark_b200.mlp import FusedMLP
    dev = cuda_device
    g = load_golden("mlp_marl_actor0")
    sd = {k[2:].replace("__", "."): v for k, v in g.items() if k.startswith("w_")}
    fused = FusedMLP.from_marl_state_dict(sd, "act.action_out.fc_mean", dev)
    x = g["x"].to(dev)
    y = fused(x)
    torch.cuda.synchronize()
    assert y.shape == (x.shape[0], 8)
    err = _rowmax_err(y.cpu(), g["mean"])
    print("MARL actor mean: max row-relative error vs fp32 reference %.3g" % err)
    assert err <= 2e-2
    # bf16-operand emulation with fp32 LayerNorm / ELU, as the kernels do.  With the deferred LayerNorm (the default) the
    # activation is rounded to bf16 BEFORE its LayerNorm, which is taken of the rounded values and applied through weights
    # with gamma folded in (include/mmb.h, `ln_in_stats`): emulated step by step
    sdd = {k: v.to(dev) for k, v in sd.items()}
    with torch.no_grad():
        h = _bf(F.layer_norm(x, (46,), sdd["base.feature_norm.weight"], sdd["base.feature_norm.bias"]))
        z = F.linear(h, _bf(sdd["base.mlp.fc1.0.weight"]), sdd["base.mlp.fc1.0.bias"])
        for pre, nxt in (("base.mlp.fc1", "base.mlp.fc2.0"), ("base.mlp.fc2.0", "base.mlp.fc2.1"), ("base.mlp.fc2.1", "act.action_out.fc_mean")):
            e = _bf(F.elu(z))
            gamma, beta = sdd[pre + ".2.weight"], sdd[pre + ".2.bias"]
            w = sdd[nxt + (".0.weight" if nxt.startswith("base") else ".weight")]
            b = sdd[nxt + (".0.bias" if nxt.startswith("base") else ".bias")]
            mean, var = e.mean(dim=1, keepdim=True), e.var(dim=1, unbiased=False, keepdim=True)
            wf = _bf(w * gamma[None, :])
            z = torch.rsqrt(var + 1e-5) * (F.linear(e, wf) - mean * wf.sum(dim=1)[None, :]) + (b + w @ beta)[None, :]
        emu = z
    assert _rowmax_err(y, emu) <= 3e-3, _rowmax_err(y, emu)


def test_mlp_layer_rejects_bad_arguments(cuda_device):
    from massive_marl_benchmark_b200 import _lib as L
    p = L.MlpLayerParams()
    assert L.lib().mmb_mlp_layer(p, None) == -1
    assert L.lib().mmb_mlp_layer(None, None) == -1
    assert L.lib().mmb_ln_cast(None, 1, 128, 4, 64, None, None, 0.0, 0, None, None) == -1


def test_ppo_act_interface_matches_reference_semantics(cuda_device):
    """`act()` returns what module.py:73-87 returns: shapes, the sigma^2 scale quirk, and log-probs that equal
    torch.distributions.MultivariateNormal's for the sampled actions."""
    from massive_marl_benchmark_b200.mlp import PPOActorCriticForward
    from torch.distributions import MultivariateNormal
    dev = cuda_device
    gen = torch.Generator().manual_seed(0)

    class AC(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.asymmetric = False
            self.actor = _ppo_net(60, [256, 128], 8, 0.01, gen)
            self.critic = _ppo_net(60, [256, 128], 1, 1.0, gen)
            self.log_std = torch.nn.Parameter(torch.log(torch.tensor(0.8)) * torch.ones(8))

    ac = AC().to(dev)
    fwd = PPOActorCriticForward(ac, dev)
    obs = torch.randn(300, 60, device=dev)
    torch.manual_seed(1)
    actions, logp, value, mean, log_std = fwd.act(obs, None)
    assert actions.shape == (300, 8) and logp.shape == (300,) and value.shape == (300, 1) and log_std.shape == (300, 8)
    assert torch.equal(log_std, ac.log_std.detach().repeat(300, 1))
    dist = MultivariateNormal(mean, scale_tril=torch.diag(ac.log_std.exp() * ac.log_std.exp()))
    assert torch.allclose(logp, dist.log_prob(actions), rtol=1e-4, atol=1e-4)
    emp_std = (actions - mean).std(dim=0)
    assert torch.allclose(emp_std, (ac.log_std.exp() ** 2).detach(), rtol=0.2)          # sigma^2, not sigma
    assert _rowmax_err(fwd.act_inference(obs), ac.actor(obs).detach()) <= 3e-2
    # actor and critic ran as one grouped launch per layer (critic head zero-padded): same numbers as on their own
    assert fwd._pair is not None
    assert torch.equal(mean, fwd.actor(obs)) and torch.equal(value, fwd.critic(obs))


def test_marl_policy_forward(cuda_device):
    from massive_marl_benchmark_b200.mlp import MarlPolicyForward
    dev = cuda_device
    g = load_golden("mlp_marl_actor0")
    sd = {k[2:].replace("__", "."): v for k, v in g.items() if k.startswith("w_")}
    torch.manual_seed(0)
    critic_sd = {k: v for k, v in sd.items() if k.startswith("base.mlp")}
    critic_sd.update({"base.feature_norm.weight": torch.ones(388), "base.feature_norm.bias": torch.zeros(388),
                      "base.mlp.fc1.0.weight": torch.randn(512, 388) * 0.05, "v_out.weight": torch.randn(1, 512) * 0.05,
                      "v_out.bias": torch.zeros(1)})
    pol = MarlPolicyForward(sd, critic_sd, device=dev)
    obs = g["x"].to(dev); share = torch.randn(obs.shape[0], 388, device=dev)
    values, actions, logp = pol.get_actions(share, obs)
    assert values.shape == (obs.shape[0], 1) and actions.shape == (obs.shape[0], 8) and logp.shape == (obs.shape[0], 8)
    v2, a_det, _ = pol.get_actions(share, obs, deterministic=True)
    assert _rowmax_err(a_det.cpu(), g["mean"]) <= 2e-2 and torch.equal(values, v2)
    std = torch.sigmoid(sd["act.action_out.log_std"] / 1.0) * 0.5
    ref_lp = torch.distributions.Normal(a_det.cpu(), std).log_prob(actions.cpu())
    assert torch.allclose(logp.cpu(), ref_lp, rtol=1e-4, atol=1e-4)


def test_grouped_mlp_equals_per_network_forward(cuda_device):
    """GroupedMLP (one launch per layer for all agents) == the same FusedMLPs run one by one, bit for bit, for both the
    MARL (LayerNorm) and the PPO (plain ELU) architectures."""
    from massive_marl_benchmark_b200.mlp import FusedMLP, GroupedMLP
    dev = cuda_device
    g = load_golden("mlp_marl_actor0")
    sd0 = {k[2:].replace("__", "."): v for k, v in g.items() if k.startswith("w_")}
    gen = torch.Generator().manual_seed(4)
    G, M = 10, 700
    sds = []
    for a in range(G):
        sds.append({k: (v + 0.01 * torch.randn(v.shape, generator=gen)) if v.dtype.is_floating_point else v for k, v in sd0.items()})
    mlps = [FusedMLP.from_marl_state_dict(sd, "act.action_out.fc_mean", dev) for sd in sds]
    xs = torch.randn(G, M, mlps[0].in_dim, generator=gen).to(dev)
    grouped = GroupedMLP(mlps)(xs)
    for a in range(G):
        assert torch.equal(grouped[a], mlps[a](xs[a])), a
    nets = [_ppo_net(60, [256, 128], 8, 0.5, gen).to(dev) for _ in range(3)]
    fm = [FusedMLP.from_sequential(n, dev) for n in nets]
    xs = torch.randn(3, 333, 60, generator=gen).to(dev)
    grouped = GroupedMLP(fm)([xs[0], xs[1], xs[2]])
    for a in range(3):
        assert torch.equal(grouped[a], fm[a](xs[a])), a


def test_marl_team_forward_equals_per_agent_policies(cuda_device):
    from massive_marl_benchmark_b200.mlp import MarlPolicyForward, MarlTeamForward
    dev = cuda_device
    g = load_golden("mlp_marl_actor0")
    sd0 = {k[2:].replace("__", "."): v for k, v in g.items() if k.startswith("w_")}
    gen = torch.Generator().manual_seed(8)
    A, N = 4, 300
    actor_sds, critic_sds = [], []
    for a in range(A):
        asd = {k: (v + 0.01 * torch.randn(v.shape, generator=gen)) if v.dtype.is_floating_point else v for k, v in sd0.items()}
        csd = {k: v for k, v in asd.items() if k.startswith("base.mlp")}
        csd.update({"base.feature_norm.weight": torch.ones(388), "base.feature_norm.bias": torch.zeros(388),
                    "base.mlp.fc1.0.weight": torch.randn(512, 388, generator=gen) * 0.05, "v_out.weight": torch.randn(1, 512, generator=gen) * 0.05,
                    "v_out.bias": torch.zeros(1)})
        actor_sds.append(asd); critic_sds.append(csd)
    team = MarlTeamForward(actor_sds, critic_sds, device=dev)
    obs = torch.randn(N, A, team.actors.in_dim, generator=gen).to(dev)        # (N, A, O) as the wrapper returns it
    share = torch.randn(N, 388, generator=gen).to(dev)
    values, actions, logp = team.get_actions(share, obs, deterministic=True)
    assert values.shape == (A, N, 1) and actions.shape == (A, N, 8) and logp.shape == (A, N, 8)
    for a in range(A):
        pol = MarlPolicyForward(actor_sds[a], critic_sds[a], device=dev)
        v, act, lp = pol.get_actions(share, obs[:, a].contiguous(), deterministic=True)
        assert torch.equal(values[a], v) and torch.equal(actions[a], act) and torch.allclose(logp[a], lp)
    # sampled: one launch for the team (per-agent std rows); the log-probs are those of Normal(mean_a, std_a) at the samples
    _, sampled, slogp = team.get_actions(share, obs)
    _, mean, _ = team.get_actions(share, obs, deterministic=True)
    for a in range(A):
        std_a = torch.sigmoid(actor_sds[a]["act.action_out.log_std"] / 1.0).to(dev) * 0.5
        ref = torch.distributions.Normal(mean[a], std_a).log_prob(sampled[a])
        assert torch.allclose(slogp[a], ref, rtol=1e-4, atol=1e-4), a
        z = (sampled[a] - mean[a]) / std_a
        assert 0.8 < float(z.std()) < 1.2 and abs(float(z.mean())) < 0.1


def test_gaussian_act_kernel(cuda_device):
    """mmb_gaussian_act: with supplied noise the actions and log-probs equal torch.distributions (PPO's sigma^2
    MultivariateNormal and MARL's per-dimension Normal); with the in-kernel Philox stream the draws are standard normal,
    reproducible for a (seed, step) and different between steps."""
    from massive_marl_benchmark_b200.mlp import gaussian_act
    from torch.distributions import MultivariateNormal, Normal
    dev = cuda_device
    gen = torch.Generator().manual_seed(0)
    M, A = 5000, 80
    mean = torch.randn(M, A, generator=gen).to(dev)
    std = (0.3 + torch.rand(A, generator=gen)).to(dev)
    z = torch.randn(M, A, generator=gen).to(dev)
    act, lp = gaussian_act(mean, std, noise=z)
    assert torch.allclose(act, mean + z * std, rtol=1e-6, atol=1e-6)
    ref = MultivariateNormal(mean, scale_tril=torch.diag(std)).log_prob(act)
    assert torch.allclose(lp, ref, rtol=1e-4, atol=2e-3)
    act2, lp2 = gaussian_act(mean, std, noise=z, per_dim=True)
    assert torch.equal(act2, act) and torch.allclose(lp2, Normal(mean, std).log_prob(act), rtol=1e-4, atol=1e-4)
    det, _ = gaussian_act(mean, std, deterministic=True)
    assert torch.equal(det, mean)
    # the broadcast output of PPO's act() (`log_std.repeat(N, 1)`, module.py:87) from the same launch, also per std group
    log_std = torch.randn(A, generator=gen).to(dev)
    act3, lp3, sig = gaussian_act(mean, std, noise=z, sigma_src=log_std)
    assert torch.equal(act3, act) and torch.equal(lp3, lp) and torch.equal(sig, log_std.repeat(M, 1))
    std_g = (0.3 + torch.rand(5, A, generator=gen)).to(dev); src_g = torch.randn(5, A, generator=gen).to(dev)
    act4, _, sig_g = gaussian_act(mean, std_g, noise=z, std_group_rows=1000, sigma_src=src_g)
    assert torch.equal(sig_g, src_g.repeat_interleave(1000, dim=0))
    assert torch.allclose(act4, mean + z * std_g.repeat_interleave(1000, dim=0), rtol=1e-6, atol=1e-6)
    a1, _ = gaussian_act(mean, std, seed=7, step=1)
    a1b, _ = gaussian_act(mean, std, seed=7, step=1)
    a2, _ = gaussian_act(mean, std, seed=7, step=2)
    assert torch.equal(a1, a1b) and not torch.equal(a1, a2)
    zz = (a1 - mean) / std
    assert abs(float(zz.mean())) < 0.01 and abs(float(zz.std()) - 1.0) < 0.01
    assert abs(float((zz ** 3).mean())) < 0.03 and abs(float((zz ** 4).mean()) - 3.0) < 0.1     # skewness, kurtosis
    rows = zz[:, :40].reshape(-1); nxt = zz[:, 1:41].reshape(-1)
    assert abs(float((rows * nxt).mean())) < 0.01                                                  # neighbours uncorrelated


def test_large_batch_takes_persistent_kernel_and_matches_row_blocks(cuda_device):
    """At M = 16384 + 100 a layer has more tiles than SMs and runs on the persistent, epilogue-overlapped kernel (two
    accumulators in tensor memory).  A row's result does not depend on the batch it is in (same K order, same epilogue),
    so the large-batch output must equal, bit for bit, the outputs of the same rows pushed through in blocks of 4096
    (which take the one-tile-per-CTA kernel) - and agree with torch fp32 to bf16 operand precision."""
    from massive_marl_benchmark_b200.mlp import FusedMLP
    dev = cuda_device
    torch.backends.cuda.matmul.allow_tf32 = False
    gen = torch.Generator().manual_seed(123)
    net = _ppo_net(388, [1024, 1024, 512], 80, 0.5, gen).to(dev)
    M = 16384 + 100
    x = torch.clamp(torch.randn(M, 388, generator=gen) * 2.0, -5, 5).to(dev)
    fused = FusedMLP.from_sequential(net, dev)
    y = fused(x).clone()
    y2 = fused(x).clone()                      # accumulators / barriers are reusable across launches
    blocks = torch.cat([fused(x[i:i + 4096]).clone() for i in range(0, M, 4096)])
    torch.cuda.synchronize()
    assert torch.equal(y, y2) and torch.equal(y, blocks)
    with torch.no_grad():
        assert _rowmax_err(y, net(x)) <= 3e-2


def test_forward_follows_the_parameters_across_optimizer_steps(cuda_device):
    """The rollout forward must act with the CURRENT policy: after `optimizer.step()` (in-place parameter update) the
    kernel-side bf16 copies are re-cast automatically (version counters), in place (same buffers); `refresh_from` re-binds
    to another module of the same architecture.  PPO pair (actor + zero-padded critic head) and a MARL state dict."""
    from massive_marl_benchmark_b200.mlp import MarlPolicyForward, PPOActorCriticForward
    dev = cuda_device
    gen = torch.Generator().manual_seed(0)

    class AC(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.asymmetric = False
            self.actor = _ppo_net(60, [256, 128], 8, 0.5, gen)
            self.critic = _ppo_net(60, [256, 128], 1, 1.0, gen)
            self.log_std = torch.nn.Parameter(torch.log(torch.tensor(0.8)) * torch.ones(8))

    ac = AC().to(dev)
    fwd = PPOActorCriticForward(ac, dev)
    obs = torch.randn(300, 60, device=dev)
    ptr_before = fwd.actor.layers[0].w.data_ptr()
    _, _, v0, m0, _ = fwd.act(obs, None)
    opt = torch.optim.Adam(ac.parameters(), lr=5e-2)
    for _ in range(3):
        opt.zero_grad()
        (ac.actor(obs).pow(2).mean() + ac.critic(obs).pow(2).mean() + ac.log_std.sum()).backward()
        opt.step()
    _, _, v1, m1, ls1 = fwd.act(obs, None)
    with torch.no_grad():
        assert _rowmax_err(m1, ac.actor(obs)) <= 3e-2 and _rowmax_err(v1, ac.critic(obs)) <= 3e-2
        assert _rowmax_err(m0, ac.actor(obs)) > 0.1, "the optimizer steps were meant to move the policy visibly"
        assert torch.equal(ls1[0], ac.log_std.detach())
    assert fwd.actor.layers[0].w.data_ptr() == ptr_before, "refresh must be in place"
    assert torch.equal(m1, fwd.actor(obs)) and torch.equal(v1, fwd.critic(obs))
    ac2 = AC().to(dev)                                   # another module object of the same architecture
    fwd.refresh_from(ac2)
    _, _, v2, m2, _ = fwd.act(obs, None)
    with torch.no_grad():
        assert _rowmax_err(m2, ac2.actor(obs)) <= 3e-2 and _rowmax_err(v2, ac2.critic(obs)) <= 3e-2
    # MARL: state dicts of live modules share storage and version counters with the parameters
    g = load_golden("mlp_marl_actor0")
    sd = {k[2:].replace("__", "."): torch.nn.Parameter(v.to(dev)) if v.dtype.is_floating_point else v for k, v in g.items() if k.startswith("w_")}
    critic_sd = {k: v for k, v in sd.items() if k.startswith("base.mlp")}
    critic_sd.update({"base.feature_norm.weight": torch.ones(388, device=dev), "base.feature_norm.bias": torch.zeros(388, device=dev),
                      "base.mlp.fc1.0.weight": torch.randn(512, 388, device=dev) * 0.05, "v_out.weight": torch.randn(1, 512, device=dev) * 0.05,
                      "v_out.bias": torch.zeros(1, device=dev)})
    pol = MarlPolicyForward(sd, critic_sd, device=dev)
    o = g["x"].to(dev); share = torch.randn(o.shape[0], 388, device=dev)
    _, a0, _ = pol.get_actions(share, o, deterministic=True)
    with torch.no_grad():
        sd["act.action_out.fc_mean.bias"].add_(0.25)         # what an optimizer step does
        sd["act.action_out.log_std"].add_(0.5)
    _, a1, _ = pol.get_actions(share, o, deterministic=True)
    assert torch.allclose(a1, a0 + 0.25, atol=1e-6)
    assert torch.allclose(pol.std, torch.sigmoid(sd["act.action_out.log_std"].detach()) * 0.5)


@pytest.mark.parametrize("M", [1, 130, 4096, 5000])
def test_single_launch_chain_equals_layer_by_layer(cuda_device, M):
    """`mmb_mlp_chain` (one launch: clusters of 4 CTAs walk all layers, cluster barrier at the layer boundaries) gives the
    numbers of the layer-by-layer launches - same bf16 operands, same epilogue; every CTA accumulates the k-blocks it
    produced itself first, so the fp32 sums differ in the last bits and a hidden activation can round to the neighbouring
    bf16 value (a few 1e-3 of the row scale at the output; bf16 ulp = 3.9e-3) - for one network and for the actor / critic
    pair, with the same bits on every repetition (a stale operand tile would show up here); geometries outside the fused
    kernel fall back silently."""
    from massive_marl_benchmark_b200 import mlp as mm
    dev = cuda_device
    gen = torch.Generator().manual_seed(M)
    torch.manual_seed(M)
    actor = _ppo_net(388, [1024, 1024, 512], 80, 0.01, gen).to(dev)
    critic = _ppo_net(388, [1024, 1024, 512], 80, 1.0, gen).to(dev)
    x = torch.clamp(torch.randn(M, 388, generator=gen) * 2.0, -5, 5).to(dev)
    outs = {}
    for chain in (True, False):
        mm._CHAIN_ENABLED = chain
        fa, fc = mm.FusedMLP.from_sequential(actor, dev), mm.FusedMLP.from_sequential(critic, dev)
        y = fa(x)
        pair = mm.GroupedMLP([fa, fc])([x, x])
        torch.cuda.synchronize()
        assert fa.__dict__.get("_chain_ok") is (True if chain else None)
        outs[chain] = (y.clone(), pair.clone())
    mm._CHAIN_ENABLED = True
    assert _rowmax_err(outs[True][0], outs[False][0]) <= 5e-3, _rowmax_err(outs[True][0], outs[False][0])
    assert _rowmax_err(outs[True][1][0], outs[False][1][0]) <= 5e-3 and _rowmax_err(outs[True][1][1], outs[False][1][1]) <= 5e-3
    fa = mm.FusedMLP.from_sequential(actor, dev)
    first = fa(x).clone()
    for _ in range(20):                                   # identical bits on every repetition
        assert torch.equal(fa(x), first)
    with torch.no_grad():
        assert _rowmax_err(outs[True][0], actor(x)) <= 3e-2
    # a geometry the fused kernel does not take (hidden width 96): falls back, same interface
    small = _ppo_net(60, [96, 96], 8, 0.5, gen).to(dev)
    fs = mm.FusedMLP.from_sequential(small, dev)
    xs = torch.randn(M, 60, generator=gen).to(dev)
    ys = fs(xs)
    assert fs._chain_ok is False
    with torch.no_grad():
        assert _rowmax_err(ys, small(xs)) <= 3e-2


@pytest.mark.parametrize("M", [1, 130, 4096])
def test_tf32_chain(cuda_device, M):
    """`FusedMLP.forward_tf32`: the single-launch chain with kind::tf32 MMAs on the fp32 observations and the LIVE fp32
    nn.Linear weights (no casts, no copies).  Against torch's fp32 SGEMM forward (the reference's path, allow_tf32 off) the
    error is well below the bf16 path's on the same inputs, and an in-place optimiser step is seen without
    any refresh."""
    from massive_marl_benchmark_b200 import mlp as mm
    dev = cuda_device
    gen = torch.Generator().manual_seed(100 + M)
    torch.manual_seed(M)
    torch.backends.cuda.matmul.allow_tf32 = False
    actor = _ppo_net(388, [1024, 1024, 512], 80, 1.0, gen).to(dev)
    x = torch.clamp(torch.randn(M, 388, generator=gen) * 2.0, -5, 5).to(dev)
    f = mm.FusedMLP.from_sequential(actor, dev)
    with torch.no_grad():
        ref = actor(x)
    e_tf32 = _rowmax_err(f.forward_tf32(x), ref)
    e_bf16 = _rowmax_err(f(x), ref)
    print("tf32 %.2e  bf16 %.2e" % (e_tf32, e_bf16))
    assert e_tf32 <= 5e-3, e_tf32            # the tensor core truncates fp32 operands to tf32 (weights and observations go in as they are)
    assert e_tf32 < 0.6 * e_bf16
    first = f.forward_tf32(x).clone()
    for _ in range(10):
        assert torch.equal(f.forward_tf32(x), first)
    with torch.no_grad():
        for prm in actor.parameters():
            prm.add_(0.01 * torch.randn(prm.shape, generator=gen).to(dev))
        ref2 = actor(x)
    assert _rowmax_err(f.forward_tf32(x), ref2) <= 5e-3


@pytest.mark.parametrize("M", [1, 130, 4096, 5000])
def test_dual_network_chain_equals_single_network_chains(cuda_device, M):
    """`mlp_chain_duo_kernel` (actor and critic on the same observations: one CTA walks both networks, alternating layer by
    layer, the fp32 rows cast once) accumulates every network's k-blocks in the order the single-network chain does, so the
    pair's outputs equal the two single launches BIT FOR BIT - on every repetition (a stale operand tile, a staging buffer
    reused too early or an accumulator overwritten under its epilogue would show up here) - and no bounded wait expired."""
    import ctypes as C
    from massive_marl_benchmark_b200 import _lib as L
    from massive_marl_benchmark_b200 import mlp as mm
    dev = cuda_device
    gen = torch.Generator().manual_seed(100 + M)
    actor = _ppo_net(388, [1024, 1024, 512], 80, 0.01, gen).to(dev)
    critic = _ppo_net(388, [1024, 1024, 512], 80, 1.0, gen).to(dev)
    x = torch.clamp(torch.randn(M, 388, generator=gen) * 2.0, -5, 5).to(dev)
    fa, fc = mm.FusedMLP.from_sequential(actor, dev), mm.FusedMLP.from_sequential(critic, dev)
    ya, yc = fa(x).clone(), fc(x).clone()
    assert fa._chain_ok is True
    pair = mm.GroupedMLP([fa, fc])
    st = (C.c_uint32 * 4)()
    L.check(L.lib().mmb_mlp_debug_status(st), "mmb_mlp_debug_status")      # (clears the record)
    for rep in range(10):
        out = pair([x, x])
        assert torch.equal(out[0], ya) and torch.equal(out[1], yc), rep
    torch.cuda.synchronize()
    L.check(L.lib().mmb_mlp_debug_status(st), "mmb_mlp_debug_status")
    assert list(st) == [0, 0, 0, 0], [hex(v) for v in st]
    with torch.no_grad():
        assert _rowmax_err(out[0], actor(x)) <= 3e-2 and _rowmax_err(out[1], critic(x)) <= 3e-2
    # a three-layer geometry whose slices are narrower (256-wide hidden layers: 64 columns per CTA, one epilogue half)
    a2 = _ppo_net(60, [256, 256], 8, 0.05, gen).to(dev)
    c2 = _ppo_net(60, [256, 256], 8, 1.0, gen).to(dev)
    x2 = torch.randn(M, 60, generator=gen).to(dev)
    f2a, f2c = mm.FusedMLP.from_sequential(a2, dev), mm.FusedMLP.from_sequential(c2, dev)
    y2a, y2c = f2a(x2).clone(), f2c(x2).clone()
    out2 = mm.GroupedMLP([f2a, f2c])([x2, x2])
    assert torch.equal(out2[0], y2a) and torch.equal(out2[1], y2c)


def test_deferred_layernorm_matches_the_in_epilogue_layernorm(cuda_device):
    """MARL trunks with the LayerNorms applied by the CONSUMING layer (gamma folded into its weights, mean / rstd from the
    producer's per-chunk partial sums: include/mmb.h `ln_in_stats`) against the first implementation (LayerNorm inside the
    producing layer's epilogue, n_tile = N) and against the fp32 torch trunk: same numbers to bf16 operand precision, for
    the actor (8-wide head, TMA-staged output) and the critic (1-wide head, row stores), single network and grouped, and
    bit-identical between the single and the grouped launch (different tile widths)."""
    from massive_marl_benchmark_b200 import mlp as mm
    dev = cuda_device
    g = load_golden("mlp_marl_actor0")
    sd = {k[2:].replace("__", "."): v.clone() for k, v in g.items() if k.startswith("w_")}
    gen = torch.Generator().manual_seed(21)
    for k in list(sd):                                   # non-trivial LayerNorm parameters
        if k.endswith(".2.weight") or k == "base.feature_norm.weight":
            sd[k] = 1.0 + 0.2 * torch.randn(sd[k].shape, generator=gen)
        if k.endswith(".2.bias") or k == "base.feature_norm.bias":
            sd[k] = 0.1 * torch.randn(sd[k].shape, generator=gen)
    csd = {k: v for k, v in sd.items() if k.startswith("base.")}
    csd.update({"v_out.weight": torch.randn(1, 512, generator=gen) * 0.05, "v_out.bias": torch.tensor([0.3])})

    def torch_trunk(d, head, x):
        h = torch.nn.functional.layer_norm(x, (x.shape[1],), d["base.feature_norm.weight"], d["base.feature_norm.bias"], 1e-5)
        names = ["base.mlp.fc1"] + ["base.mlp.fc2.%d" % i for i in range(8) if "base.mlp.fc2.%d.0.weight" % i in d]
        for n in names:
            h = torch.nn.functional.elu(h @ d[n + ".0.weight"].T + d[n + ".0.bias"])
            h = torch.nn.functional.layer_norm(h, (h.shape[1],), d[n + ".2.weight"], d[n + ".2.bias"], 1e-5)
        return h @ d[head + ".weight"].T + d[head + ".bias"]

    for M in (1, 130, 700, 4096):
        x = torch.randn(M, sd["base.feature_norm.weight"].shape[0], generator=gen)
        for d, head in ((sd, "act.action_out.fc_mean"), (csd, "v_out")):
            want = torch_trunk(d, head, x)
            outs = {}
            for defer in (True, False):
                mm._DEFER_LN = defer
                f = mm.FusedMLP.from_marl_state_dict(d, head, dev)
                assert any(l.fold is not None for l in f.layers) is defer
                outs[defer] = f(x.to(dev)).cpu()
                if defer:
                    grouped = mm.GroupedMLP([f, mm.FusedMLP.from_marl_state_dict(d, head, dev)])([x.to(dev), x.to(dev)])
                    assert torch.equal(grouped[0].cpu(), outs[True]) and torch.equal(grouped[1].cpu(), outs[True]), (M, head)
            mm._DEFER_LN = True
            e_def, e_epi, e_x = _rowmax_err(outs[True], want), _rowmax_err(outs[False], want), _rowmax_err(outs[True], outs[False])
            print("M %d %s: deferred %.4f, in-epilogue %.4f vs fp32; deferred vs in-epilogue %.4f" % (M, head, e_def, e_epi, e_x))
            # worst row of up to 4096 with random LayerNorm scales: both bf16 paths sit at 1-4 % of a row's largest output, and
            # the deferred form is no further from fp32 than the in-epilogue form (measured: 0.0255 vs 0.0239, 0.0381 vs 0.0357
            # at M = 4096; the two differ from each other by 0.4-1.2 %)
            assert e_def <= 5e-2 and e_epi <= 5e-2 and e_x <= 5e-2, (M, head, e_def, e_epi, e_x)
            if M >= 130:
                assert e_def <= 1.25 * e_epi + 2e-3, (M, head, e_def, e_epi)
    # the folded weights follow the live parameters: an in-place change of a LayerNorm's gamma re-folds the next layer
    mm._DEFER_LN = True
    live = {k: v.clone().to(dev) for k, v in sd.items()}
    f = mm.FusedMLP.from_marl_state_dict(live, "act.action_out.fc_mean", dev)
    x = torch.randn(64, live["base.feature_norm.weight"].shape[0], generator=gen).to(dev)
    y0 = f(x).clone()
    live["base.mlp.fc1.2.weight"].mul_(1.5)
    y1 = f(x)
    cpu = {k: v.cpu() for k, v in live.items()}
    assert not torch.equal(y0, y1) and _rowmax_err(y1.cpu(), torch_trunk(cpu, "act.action_out.fc_mean", x.cpu())) <= 2e-2


def test_dual_network_chain_small_batches_forced(cuda_device):
    """Below ~2300 rows the side-by-side launch is chosen, so the dual-network kernel's handling of partial row blocks
    (M = 1, 130, 300: rows beyond M never stored, TMA-clipped) is exercised in a child process with MMB_MLP_DUO=2 (the switch
    is read once per process): the pair equals the two single-network launches bit for bit there too."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, torch
sys.path.insert(0, %r)
from massive_marl_benchmark_b200 import mlp as mm
import ctypes as C
from massive_marl_benchmark_b200 import _lib as L
dev = torch.device("cuda:0")
def net(dims):
    mods = []
    for i in range(len(dims) - 1):
        mods.append(torch.nn.Linear(dims[i], dims[i + 1]))
        if i < len(dims) - 2:
            mods.append(torch.nn.ELU())
    return torch.nn.Sequential(*mods).to(dev)
torch.manual_seed(3)
for dims in ([388, 1024, 1024, 512, 80], [60, 256, 256, 8]):
    a, c = net(dims), net(dims)
    fa, fc = mm.FusedMLP.from_sequential(a, dev), mm.FusedMLP.from_sequential(c, dev)
    pair = mm.GroupedMLP([fa, fc])
    for M in (1, 130, 300):
        x = torch.randn(M, dims[0], device=dev)
        ya, yc = fa(x).clone(), fc(x).clone()
        for rep in range(3):
            out = pair([x, x])
            assert torch.equal(out[0], ya) and torch.equal(out[1], yc), (dims, M, rep)
st = (C.c_uint32 * 4)()
L.lib().mmb_mlp_debug_status(st)
assert list(st) == [0, 0, 0, 0], [hex(v) for v in st]
print("forced duo ok")
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=dict(os.environ, MMB_MLP_DUO="2"))
    assert r.returncode == 0 and "forced duo ok" in r.stdout, (r.stdout[-500:], r.stderr[-1500:])
