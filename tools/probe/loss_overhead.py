"""Where the host time of the `ppo_loss` autograd wrapper goes (B = 16384, A = 80): the raw launch, the Function's forward,
the backward, and the floor an empty autograd.Function with the same inputs / outputs costs in this torch build."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from massive_marl_benchmark_b200.ppo_loss import ppo_loss, ppo_loss_raw  # noqa: E402

dev = torch.device("cuda:0")
B, A = 16384, 80
g = torch.Generator(device=dev).manual_seed(1)
r = lambda *s: torch.randn(*s, generator=g, device=dev)  # noqa: E731
log_std = (r(A) * 0.1 - 0.3).requires_grad_(True)
old_mu = r(B, A) * 0.5
mu = (old_mu + 0.01 * r(B, A)).requires_grad_(True)
actions = old_mu + 0.5 * r(B, A)
old_sigma = log_std.detach().repeat(B, 1).contiguous()
old_logp, adv, tv = r(B, 1), r(B, 1), r(B, 1)
value = (tv + 0.3 * r(B, 1)).requires_grad_(True)
ret = tv + 0.5 * r(B, 1)


class Floor(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, log_std, value):
        ctx.save_for_backward(mu, log_std, value)
        return mu.new_zeros(())

    @staticmethod
    def backward(ctx, gl):
        return ctx.saved_tensors


def host_and_device(fn, iters=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / iters * 1e6
    torch.cuda.synchronize()
    return {"host_us": round(host, 1), "device_us": round(e0.elapsed_time(e1) / iters * 1e3, 1)}


def fwd():
    return ppo_loss(mu, log_std, value, actions, old_logp, adv, tv, ret, old_mu, old_sigma)


def fwd_bwd():
    for t in (mu, log_std, value):
        t.grad = None
    fwd().loss.backward()


def floor():
    for t in (mu, log_std, value):
        t.grad = None
    Floor.apply(mu, log_std, value).backward()


out = {"raw_launch": host_and_device(lambda: ppo_loss_raw(mu, log_std, value, actions, old_logp, adv, tv, ret, old_mu, old_sigma)),
       "function_forward": host_and_device(fwd), "forward_backward": host_and_device(fwd_bwd),
       "empty_function_forward_backward": host_and_device(floor)}
# the MAPPO loss wrapper (B = 16384, A = 8): in-kernel finalisation against the fp64 sums reduced by torch ops
from massive_marl_benchmark_b200.mappo_loss import mappo_loss, mappo_loss_raw  # noqa: E402

Am = 8
m_mean = (r(B, Am) * 0.3).requires_grad_(True)
m_log_std = torch.zeros(Am, device=dev, requires_grad=True)
m_actions = m_mean.detach() + 0.3 * r(B, Am)
m_old_logp = r(B, Am) * 0.1 - 1.0
m_values = r(B, 1).requires_grad_(True)
m_vp, m_ret, m_adv = r(B, 1), r(B, 1), r(B, 1)


def m_args():
    return (m_mean, torch.sigmoid(m_log_std) * 0.5, m_values, m_actions, m_old_logp, m_adv, m_vp, m_ret)


def m_fwd_bwd():
    for t in (m_mean, m_log_std, m_values):
        t.grad = None
    o = mappo_loss(*m_args())
    (o.policy_loss - o.dist_entropy * 0.01).backward()
    o.value_loss.backward()


out["mappo_raw_finalised"] = host_and_device(lambda: mappo_loss_raw(*m_args(), finalise=True))
out["mappo_raw_torch_reduced"] = host_and_device(lambda: mappo_loss_raw(*m_args(), finalise=False))
out["mappo_forward"] = host_and_device(lambda: mappo_loss(*m_args()))
out["mappo_forward_two_backwards"] = host_and_device(m_fwd_bwd)
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/loss_overhead.json", "w"), indent=1)
