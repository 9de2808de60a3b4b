"""Throughput of the tcgen05 MLP forward vs torch (fp32 SGEMM = the reference's path, and bf16) on one GPU.
PPO ActorCritic at N = 4096 (obs 388 -> 1024 -> 1024 -> 512 -> 80 / 1) and the MARL actor/critic (512-wide, LayerNorm)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from massive_marl_benchmark_b200 import _lib as L  # noqa: E402
from massive_marl_benchmark_b200.mlp import FusedMLP  # noqa: E402


def net(dims):
    mods = []
    for i in range(len(dims) - 1):
        mods.append(torch.nn.Linear(dims[i], dims[i + 1]))
        if i < len(dims) - 2:
            mods.append(torch.nn.ELU())
    return torch.nn.Sequential(*mods)


def timeit(fn, iters=50, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
out = {}
M = 4096
for name, dims in (("ppo_actor", [388, 1024, 1024, 512, 80]), ("ppo_critic", [388, 1024, 1024, 512, 1])):
    n = net(dims).to(dev)
    x = torch.randn(M, dims[0], device=dev)
    f = FusedMLP.from_sequential(n, dev)
    flops = 2 * M * sum(dims[i] * dims[i + 1] for i in range(len(dims) - 1))
    with torch.no_grad():
        t_fp32 = timeit(lambda: n(x))
        nb = n.to(torch.bfloat16); xb = x.to(torch.bfloat16)
        t_bf16 = timeit(lambda: nb(xb))
    t_ours = timeit(lambda: f(x))
    L.profile_enable(True); L.profile_collect()
    for _ in range(20):
        f(x)
    torch.cuda.synchronize(); L.profile_enable(False)
    prof = L.profile_collect()
    out[name] = {"M": M, "dims": dims, "gflop": flops / 1e9, "ours_ms": t_ours, "torch_fp32_ms": t_fp32, "torch_bf16_ms": t_bf16,
                 "ours_tflops": flops / t_ours / 1e9, "torch_fp32_tflops": flops / t_fp32 / 1e9,
                 "kernel_ms": {k: v[0] / v[1] for k, v in prof.items()}, "kernel_launches": {k: v[1] / 20 for k, v in prof.items()}}
# ---- the tf32 variant of the single-launch chain: error and rate beside the bf16 path, both against torch fp32 ----
dims = [388, 1024, 1024, 512, 80]
n32 = net(dims).to(dev)
x = torch.clamp(torch.randn(M, dims[0], device=dev) * 2.0, -5, 5)
f = FusedMLP.from_sequential(n32, dev)
with torch.no_grad():
    ref = n32(x)


def rowmax_err(a, b):
    floor = b.pow(2).mean().sqrt().clamp(min=1e-6)
    return float(((a - b).abs().amax(dim=1) / torch.maximum(b.abs().amax(dim=1), floor)).max())


flops = 2 * M * sum(dims[i] * dims[i + 1] for i in range(len(dims) - 1))
t_tf32, t_bf16c = timeit(lambda: f.forward_tf32(x)), timeit(lambda: f(x))
out["ppo_actor_tf32"] = {"M": M, "dims": dims, "tf32_ms": t_tf32, "tf32_tflops": flops / t_tf32 / 1e9, "bf16_ms": t_bf16c,
                         "rowmax_err_vs_torch_fp32": {"tf32": rowmax_err(f.forward_tf32(x), ref), "bf16": rowmax_err(f(x), ref)},
                         "max_abs_err_vs_torch_fp32": {"tf32": float((f.forward_tf32(x) - ref).abs().max()), "bf16": float((f(x) - ref).abs().max())},
                         "output_rms": float(ref.pow(2).mean().sqrt())}
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_mlp.json", "w"), indent=1)

# ---- PPO act(): actor + critic as one grouped launch per layer vs one network after the other -----------------------------
from massive_marl_benchmark_b200.mlp import PPOActorCriticForward  # noqa: E402


class _AC(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.asymmetric = False
        self.actor, self.critic = net([388, 1024, 1024, 512, 80]), net([388, 1024, 1024, 512, 1])
        self.log_std = torch.nn.Parameter(torch.zeros(80))


ac = _AC().to(dev)
fwd = PPOActorCriticForward(ac, dev)
xo = torch.randn(M, 388, device=dev)
t_pair = timeit(lambda: fwd._mean_value(xo, None))
t_sep = timeit(lambda: (fwd.actor(xo), fwd.critic(xo)))
out["ppo_actor_critic"] = {"M": M, "grouped_ms": t_pair, "one_by_one_ms": t_sep, "gflop": 32.6, "grouped_tflops": 32.6 / t_pair}
print(json.dumps(out["ppo_actor_critic"]))

# ---- MARL: ten per-agent actors (46 -> 512 -> 512 -> 512 -> 8 with LayerNorms) and critics (388 -> ... -> 1) at M = 4096,
# one network after the other (the reference's loop, runner.py:205-217) vs one grouped launch per layer -----------------
from massive_marl_benchmark_b200.mlp import GroupedMLP  # noqa: E402


def marl_sd(in_dim, out_dim, head, gen):
    sd = {"base.feature_norm.weight": torch.ones(in_dim), "base.feature_norm.bias": torch.zeros(in_dim)}
    dims = [in_dim, 512, 512, 512]
    names = ["base.mlp.fc1", "base.mlp.fc2.0", "base.mlp.fc2.1"]
    for i, n in enumerate(names):
        sd[n + ".0.weight"] = torch.randn(512, dims[i], generator=gen) * (1.0 / dims[i] ** 0.5)
        sd[n + ".0.bias"] = torch.zeros(512)
        sd[n + ".2.weight"] = torch.ones(512); sd[n + ".2.bias"] = torch.zeros(512)
    sd[head + ".weight"] = torch.randn(out_dim, 512, generator=gen) * 0.05
    sd[head + ".bias"] = torch.zeros(out_dim)
    return sd


gen = torch.Generator().manual_seed(0)
for name, in_dim, out_dim, head in (("marl_actors_x10", 46, 8, "act.action_out.fc_mean"), ("marl_critics_x10", 388, 1, "v_out")):
    mlps = [FusedMLP.from_marl_state_dict(marl_sd(in_dim, out_dim, head, gen), head, dev) for _ in range(10)]
    xs = torch.randn(10, M, in_dim, device=dev)
    grp = GroupedMLP(mlps)
    t_seq = timeit(lambda: [m(xs[i]) for i, m in enumerate(mlps)], iters=20, warm=5)
    t_grp = timeit(lambda: grp(xs), iters=20, warm=5)
    flops = 10 * 2 * M * (in_dim * 512 + 2 * 512 * 512 + 512 * out_dim)
    out[name] = {"M": M, "agents": 10, "gflop": flops / 1e9, "one_by_one_ms": t_seq, "grouped_ms": t_grp,
                 "grouped_tflops": flops / t_grp / 1e9, "launches_one_by_one": 50, "launches_grouped": 5}
print(json.dumps({k: out[k] for k in ("marl_actors_x10", "marl_critics_x10")}, indent=1))
json.dump(out, open("gpurun_out/bench_mlp.json", "w"), indent=1)
