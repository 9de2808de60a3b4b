"""GPU: the fused PPO minibatch loss (`mmb_ppo_loss` through massive_marl_benchmark_b200.ppo_loss) against the oracle
(oracle/ppo_loss_oracle.py: the reference's statements + torch autograd on the CPU) and the golden fixture generated from
the reference itself.

Tolerance: value-side quantities 1e-5 relative.  Everything downstream of `ratio = exp(logp - old_logp)` inherits the
conditioning of the reference's own fp32 evaluation: an ABSOLUTE rounding error of the log-probability becomes the same
RELATIVE error of the ratio, and one ulp of a log-probability of magnitude ~100 (TenAnt, 80 action dims) is 7.6e-6, so two
correct fp32 evaluations with different summation orders differ by a few of those; these quantities are compared at
max(1e-5, 4 ulp of the largest |logp|) of the tensor's scale.  The KL estimate sums A terms that are each a difference of
numbers near 0.5 cancelling to ~1e-3, and exp(log_std) is shared by all rows, so a 1-2 ulp difference between the CPU's
and the GPU's expf does not average out over the batch: against the CPU oracle 1e-5 relative + 2.4e-7 * sqrt(A)
absolute; against the SAME oracle statements evaluated by torch on the GPU (same expf) 1e-5 relative + 5e-8."""
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _close(a, b, rel, what, floor=0.0):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    scale = float(b.abs().max()) + 1e-12
    err = float((a - b).abs().max())
    assert err <= rel * scale + floor, "%s: max err %.3e at scale %.3e (rel %.1e)" % (what, err, scale, rel)


def _check(dev, mb, cfg, want):
    from massive_marl_benchmark_b200.ppo_loss import ppo_loss
    d = {k: v.to(dev) for k, v in mb.items()}
    mu = d["mu"].clone().requires_grad_(True)
    log_std = d["log_std"].clone().requires_grad_(True)
    value = d["value"].clone().requires_grad_(True)
    out = ppo_loss(mu, log_std, value, d["actions"], d["old_logp"], d["advantages"], d["target_values"], d["returns"],
                   d["old_mu"], d["old_sigma"], **cfg)
    out.loss.backward()
    _close(out.logp, want["logp"], 2e-6, "logp")
    _close(out.entropy, want["entropy"], 1e-5, "entropy")
    _close(out.value_loss, want["value_loss"], 1e-5, "value_loss")
    _close(value.grad, want["grad_value"], 1e-5, "grad_value")
    rr = max(1e-5, 4 * 1.1920929e-07 * float(want["logp"].abs().max()))       # 4 ulp of the log-probability, see the header
    A = mu.shape[1]
    _close(out.kl_mean, want["kl_mean"], 1e-5, "kl_mean", floor=2.4e-7 * A ** 0.5)
    from oracle.ppo_loss_oracle import ppo_loss_oracle
    on_gpu = ppo_loss_oracle(**d, **cfg)                      # the oracle's statements with the GPU's transcendental functions
    _close(out.kl_mean, on_gpu["kl_mean"], 1e-5, "kl_mean (oracle statements on the GPU)", floor=5e-8)
    _close(mu.grad, on_gpu["grad_mu"], rr, "grad_mu (oracle statements on the GPU)")
    _close(log_std.grad, on_gpu["grad_log_std"], rr, "grad_log_std (oracle statements on the GPU)")
    # the surrogate is a mean of signed terms that largely cancel: its error scales with the terms, not with the mean
    term_scale = float((torch.exp(want["logp"] - mb["old_logp"].reshape(-1)) * mb["advantages"].reshape(-1)).abs().mean())
    _close(out.surrogate_loss, want["surrogate_loss"], rr, "surrogate_loss", floor=rr * term_scale)
    _close(out.loss, want["loss"], rr, "loss", floor=rr * term_scale)
    _close(mu.grad, want["grad_mu"], rr, "grad_mu")
    _close(log_std.grad, want["grad_log_std"], rr, "grad_log_std")
    assert value.grad.shape == value.shape and log_std.grad.shape == log_std.shape
    return out


def test_ppo_loss_against_the_reference_fixture(cuda_device):
    g = load_golden("ppo_loss")
    for tag in ("a8", "a80", "a8u"):
        mb = {k[len(tag) + 5:]: v for k, v in g.items() if k.startswith(tag + "__in_")}
        cfg = {k[len(tag) + 6:]: float(v) for k, v in g.items() if k.startswith(tag + "__cfg_")}
        cfg["use_clipped_value_loss"] = bool(cfg["use_clipped_value_loss"])
        want = {k[len(tag) + 6:]: v for k, v in g.items() if k.startswith(tag + "__out_")}
        _check(cuda_device, mb, cfg, want)


@pytest.mark.parametrize("B,A", [(1, 8), (31, 3), (1000, 8), (777, 16), (4100, 24), (513, 80), (300, 200), (16384, 80)])
def test_ppo_loss_against_the_oracle(cuda_device, B, A):
    """Ragged row counts (partial last group / warp / block), every lane-group width and columns-per-lane case."""
    from oracle.ppo_loss_oracle import ppo_loss_oracle, synthetic_minibatch
    mb = synthetic_minibatch(B, A, seed=B + A, ratio_spread=0.15 if A <= 24 else 0.05)
    for cfg in (dict(clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.0, use_clipped_value_loss=True),
                dict(clip_param=0.1, value_loss_coef=0.5, entropy_coef=0.02, use_clipped_value_loss=False)):
        want = ppo_loss_oracle(**mb, **cfg)
        _check(cuda_device, mb, cfg, want)


def test_ppo_loss_gradient_rules_on_ties_and_clip_edges(cuda_device):
    """Hand-built rows: advantage 0 (tie outside the clip range -> zero gradient), ratio exactly 1 (tie inside -> full
    gradient), ratio above / below the range with either advantage sign, value exactly at the old value (tie of the two
    value losses) and outside the value clip range on the side where the clipped loss wins (zero gradient)."""
    from oracle.ppo_loss_oracle import ppo_loss_oracle
    A = 8
    log_std = torch.full((A,), -0.2)
    std = log_std.exp() * log_std.exp()
    shifts = torch.tensor([0.0, 0.0, 0.6, 0.6, -0.6, -0.6, 0.05, -0.05])
    B = shifts.numel()
    old_mu = torch.zeros(B, A)
    actions = old_mu + 0.5 * std
    mu = old_mu + shifts[:, None] * std / A
    old_sigma = log_std.repeat(B, 1)
    old_logp = (-0.5 * (((actions - old_mu) / std) ** 2).sum(-1) - std.log().sum() - 0.5 * A * 1.8378770664093453).view(B, 1)
    advantages = torch.tensor([0.0, 1.0, 1.0, -1.0, 1.0, -1.0, 0.0, 2.0]).view(B, 1)
    target_values = torch.zeros(B, 1)
    value = torch.tensor([0.0, 0.1, 0.5, 0.5, -0.5, -0.5, 0.2, -0.2]).view(B, 1)
    returns = torch.tensor([1.0, 1.0, 1.0, -1.0, 1.0, -1.0, 0.0, 0.0]).view(B, 1)
    mb = dict(mu=mu, log_std=log_std, value=value, actions=actions, old_logp=old_logp, advantages=advantages,
              target_values=target_values, returns=returns, old_mu=old_mu, old_sigma=old_sigma)
    cfg = dict(clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.0, use_clipped_value_loss=True)
    want = ppo_loss_oracle(**mb, **cfg)
    assert (want["grad_mu"][0] == 0).all() and (want["grad_mu"][1] != 0).any()        # the fixture hits the intended cases
    assert (want["grad_value"] == 0).any() and (want["grad_value"] != 0).any()
    _check(cuda_device, mb, cfg, want)


def test_ppo_loss_without_kl_rows_and_argument_errors(cuda_device):
    from massive_marl_benchmark_b200 import _lib as L
    from massive_marl_benchmark_b200.ppo_loss import ppo_loss, ppo_loss_raw
    from oracle.ppo_loss_oracle import ppo_loss_oracle, synthetic_minibatch
    dev = cuda_device
    mb = synthetic_minibatch(200, 8, seed=5)
    d = {k: v.to(dev) for k, v in mb.items()}
    want = ppo_loss_oracle(**mb)
    out = ppo_loss(d["mu"], d["log_std"], d["value"], d["actions"], d["old_logp"], d["advantages"], d["target_values"],
                   d["returns"])                                   # no old_mu / old_sigma: KL estimate skipped
    assert float(out.kl_mean) == 0.0
    _close(out.loss, want["loss"], 1e-5, "loss")
    # strided mean (a column slice of a wider MLP output, as the padded grouped head produces)
    wide = torch.zeros(200, 80, device=dev)
    wide[:, :8] = d["mu"]
    sums, logp, gmu, gv = ppo_loss_raw(wide[:, :8], d["log_std"], d["value"], d["actions"], d["old_logp"], d["advantages"],
                                       d["target_values"], d["returns"], d["old_mu"], d["old_sigma"])
    _close(logp, want["logp"], 2e-6, "logp (strided mean)")
    _close(gmu, want["grad_mu"], 1e-5, "grad_mu (strided mean)")
    with pytest.raises(ValueError):
        ppo_loss(d["mu"], d["log_std"], d["value"][:-1], d["actions"], d["old_logp"], d["advantages"], d["target_values"], d["returns"])
    with pytest.raises(L.MmbError):                                # act_dim > 256 is refused by the library
        big = torch.zeros(4, 300, device=dev)
        ppo_loss_raw(big, torch.zeros(300, device=dev), torch.zeros(4, device=dev), big, torch.zeros(4, device=dev),
                     torch.zeros(4, device=dev), torch.zeros(4, device=dev), torch.zeros(4, device=dev))
    with pytest.raises(L.MmbError):
        ppo_loss_raw(mb["mu"], mb["log_std"], mb["value"], mb["actions"], mb["old_logp"], mb["advantages"],
                     mb["target_values"], mb["returns"])          # CPU tensors: no CPU path
