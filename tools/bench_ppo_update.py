"""One full PPO update (cfg/ppo/config.yaml: 5 epochs x 4 minibatches, Adam, adaptive KL) on a TenAnt-sized rollout
(4096 envs x 16 steps, obs 388, 80 actions, actor / critic 388-1024-1024-512-{80,1}):
  ours       massive_marl_benchmark_b200.ppo_update (device shuffle, fused gather, torch MLPs, fused loss kernel)
  reference  the same loop as the reference runs it on the GPU: Python-list minibatch indices from a CPU randperm, nine
             advanced-index gathers, ~40 torch kernels for distribution + loss, three host read-backs per minibatch
             (restated in oracle.ppo_loss_oracle.ppo_update_oracle, which is pinned against the reference's PPO.update)
Both use torch for the MLP forward / backward and the optimiser, so the difference is the storage-side and loss-side work
this library replaces.  Writes gpurun_out/bench_ppo_update.json.  (Measurement tool: the oracle is used here as the
reference leg being timed, never by the product path.)"""
import copy
import json
import os
import sys
import time
import types

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from massive_marl_benchmark_b200 import _lib as L  # noqa: E402
from massive_marl_benchmark_b200.ppo_update import ppo_update  # noqa: E402
from massive_marl_benchmark_b200.storage import RolloutStorage  # noqa: E402
from oracle.ppo_loss_oracle import ppo_update_oracle  # noqa: E402

dev = torch.device("cuda:0")
T, N, OBS, A = 16, 4096, 388, 80
EPOCHS, MINIBATCHES = 5, 4


class ActorCritic(nn.Module):
    def __init__(self):
        super().__init__()

        def mlp(out):
            return nn.Sequential(nn.Linear(OBS, 1024), nn.ELU(), nn.Linear(1024, 1024), nn.ELU(), nn.Linear(1024, 512), nn.ELU(),
                                 nn.Linear(512, out))
        self.actor, self.critic = mlp(A), mlp(1)
        self.log_std = nn.Parameter(torch.log(torch.tensor(0.8)) * torch.ones(A))


torch.manual_seed(0)
ac0 = ActorCritic().to(dev)
st = RolloutStorage(N, T, (OBS,), (0,), (A,), dev, "random")
g = torch.Generator(device=dev).manual_seed(1)
with torch.no_grad():
    st.observations.copy_(torch.randn(T, N, OBS, generator=g, device=dev).clamp_(-5, 5))
    flat = st.observations.view(-1, OBS)
    mu = ac0.actor(flat).view(T, N, A)
    val = ac0.critic(flat).view(T, N, 1)
    std = ac0.log_std.exp() * ac0.log_std.exp()
    act = mu + std * torch.randn(T, N, A, generator=g, device=dev)
    logp = (-0.5 * (((act - mu) / std) ** 2).sum(-1) - std.log().sum() - 0.5 * A * 1.8378770664093453).view(T, N, 1)
    st.mu.copy_(mu); st.sigma.copy_(ac0.log_std.detach().repeat(T, N, 1)); st.actions.copy_(act); st.actions_log_prob.copy_(logp)
    st.values.copy_(val + 0.1 * torch.randn(T, N, 1, generator=g, device=dev))
    st.returns.copy_(val + 0.5 * torch.randn(T, N, 1, generator=g, device=dev))
    adv = st.returns - st.values
    st.advantages.copy_((adv - adv.mean()) / (adv.std() + 1e-8))


def make(ac):
    return types.SimpleNamespace(storage=st, actor_critic=ac, optimizer=torch.optim.Adam(ac.parameters(), lr=3e-4),
                                 num_mini_batches=MINIBATCHES, num_learning_epochs=EPOCHS, clip_param=0.2, value_loss_coef=2.0,
                                 entropy_coef=0.0, use_clipped_value_loss=True, desired_kl=0.016, schedule="adaptive",
                                 step_size=3e-4, max_grad_norm=1.0, asymmetric=False)


def timed(fn, reps=3):
    fn()                                             # warm-up (cuBLAS handles, autograd graphs, lazy module loads)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def ours():
    ppo_update(make(copy.deepcopy(ac0)))


def reference():
    # storage.py:75-87: BatchSampler(SubsetRandomSampler) -> Python lists of ints, re-drawn per epoch
    orders = [torch.randperm(T * N).tolist() for _ in range(EPOCHS)]
    ppo_update_oracle(make(copy.deepcopy(ac0)), orders)


n0 = L.launch_count()
t_ours = timed(ours)
launches = (L.launch_count() - n0) / 4
t_ref = timed(reference)
out = {"transitions": T * N, "epochs": EPOCHS, "minibatches": MINIBATCHES, "ours_ms": t_ours, "reference_loop_on_gpu_ms": t_ref,
       "speedup": t_ref / t_ours, "library_launches_per_update": launches,
       "note": "wall clock around one update incl. the final synchronize; MLP forward/backward and Adam are torch in both"}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_ppo_update.json", "w"), indent=1)
