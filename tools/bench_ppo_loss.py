"""PPO minibatch loss, forward + backward down to the MLP outputs (ppo.py:266-306 minus the two MLPs): the fused
`mmb_ppo_loss` launch against the reference's torch statements + autograd on the same GPU.  TenAnt PPO minibatch
(16384 rows x 80 action dims) and OneAnt width (8).  Writes gpurun_out/bench_ppo_loss.json."""
import json
import os
import sys

import torch
from torch.distributions import MultivariateNormal

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from massive_marl_benchmark_b200 import _lib as L  # noqa: E402
from massive_marl_benchmark_b200.ppo_loss import ppo_loss, ppo_loss_raw  # noqa: E402

dev = torch.device("cuda:0")
out = {}


def timed(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def torch_loss(mu, log_std, value, actions, old_logp, adv, tv, ret, old_mu, old_sigma, clip=0.2):
    cov = torch.diag(log_std.exp() * log_std.exp())
    dist = MultivariateNormal(mu, scale_tril=cov)
    logp = dist.log_prob(actions)
    ent = dist.entropy()
    sigma = log_std.repeat(mu.shape[0], 1)
    kl = torch.sum(sigma - old_sigma + (torch.square(old_sigma.exp()) + torch.square(old_mu - mu)) / (2.0 * torch.square(sigma.exp())) - 0.5, axis=-1)
    kl_mean = torch.mean(kl)
    ratio = torch.exp(logp - torch.squeeze(old_logp))
    s = -torch.squeeze(adv) * ratio
    sc = -torch.squeeze(adv) * torch.clamp(ratio, 1.0 - clip, 1.0 + clip)
    sl = torch.max(s, sc).mean()
    vc = tv + (value - tv).clamp(-clip, clip)
    vl = torch.max((value - ret).pow(2), (vc - ret).pow(2)).mean()
    return sl + 1.0 * vl - 0.0 * ent.mean(), kl_mean


for B, A in ((16384, 80), (16384, 8), (65536, 80)):
    g = torch.Generator(device=dev).manual_seed(B + A)
    r = lambda *s: torch.randn(*s, generator=g, device=dev)  # noqa: E731
    log_std = (r(A) * 0.1 - 0.3)
    old_mu = r(B, A) * 0.5
    mu0 = old_mu + 0.01 * r(B, A)
    std = log_std.exp() ** 2
    actions = old_mu + std * r(B, A)
    old_sigma = log_std.repeat(B, 1).contiguous()
    old_logp = MultivariateNormal(old_mu, scale_tril=torch.diag(std)).log_prob(actions).view(B, 1)
    adv, tv = r(B, 1), r(B, 1)
    value0 = tv + 0.3 * r(B, 1)
    ret = tv + 0.5 * r(B, 1)

    def run_torch():
        mu = mu0.clone().requires_grad_(True); ls = log_std.clone().requires_grad_(True); v = value0.clone().requires_grad_(True)
        loss, _ = torch_loss(mu, ls, v, actions, old_logp, adv, tv, ret, old_mu, old_sigma)
        loss.backward()
        return mu.grad

    def run_ours():
        mu = mu0.clone().requires_grad_(True); ls = log_std.clone().requires_grad_(True); v = value0.clone().requires_grad_(True)
        o = ppo_loss(mu, ls, v, actions, old_logp, adv, tv, ret, old_mu, old_sigma)
        o.loss.backward()
        return mu.grad

    def run_raw():
        return ppo_loss_raw(mu0, log_std, value0, actions, old_logp, adv, tv, ret, old_mu, old_sigma)

    ga, gb = run_torch(), run_ours()
    rel = float((ga - gb).abs().max() / ga.abs().max())
    t_torch, t_ours, t_raw = timed(run_torch), timed(run_ours), timed(run_raw)
    L.profile_enable(True); L.profile_collect()
    for _ in range(20):
        run_raw()
    torch.cuda.synchronize(); L.profile_enable(False)
    tot, cnt = L.profile_collect()["ppo_loss"]
    k_ms = tot / cnt
    bytes_alg = B * (5 * A * 4 + 8 * 4)
    out["B%d_A%d" % (B, A)] = {"torch_fwd_bwd_ms": t_torch, "ours_autograd_fn_ms": t_ours, "ours_launch_plus_allocs_ms": t_raw,
                               "kernel_ms": k_ms, "algorithmic_MB": bytes_alg / 1e6, "kernel_GBps": bytes_alg / k_ms / 1e6,
                               "grad_mu_rel_diff_vs_torch": rel}
    print(B, A, out["B%d_A%d" % (B, A)], flush=True)

os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_ppo_loss.json", "w"), indent=1)
