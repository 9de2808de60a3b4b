"""The PPO rollout loop of the reference WITH the policy in it (agents/algorithms/rl/ppo/ppo.py:127-151: act -> env.step ->
add_transitions -> current_obs.copy_(next_obs), then compute_returns per rollout) on the replacement classes, TenAnt
N = 4096, horizon 16, frames resident in HBM:
  eager    VecTaskPython + PPOActorCriticForward.act + RolloutStorage.add_transitions
  graphed  GraphedVecTaskPython for the env part
  graphed_rollout  ppo_rollout.GraphedPPORollout: the whole rollout (T x [act, step, insert] + compute_returns) as ONE graph replay
and the reference's own structure on the same GPU (the torch ActorCritic module of module.py:25-107 in fp32 + the same env /
storage replacements) for the policy's share.  Host time and device time per env step.  Writes gpurun_out/bench_ppo_rollout.json."""
import json
import os
import sys
import time

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from massive_marl_benchmark_b200 import synthetic  # noqa: E402
from massive_marl_benchmark_b200.mlp import PPOActorCriticForward  # noqa: E402
from massive_marl_benchmark_b200.ppo_rollout import GraphedPPORollout  # noqa: E402
from massive_marl_benchmark_b200.providers import ReplayProvider  # noqa: E402
from massive_marl_benchmark_b200.storage import RolloutStorage  # noqa: E402
from massive_marl_benchmark_b200.tasks import TenAnt  # noqa: E402
from massive_marl_benchmark_b200.vec_task import GraphedVecTaskPython, VecTaskPython  # noqa: E402

dev = torch.device("cuda:0")
N, T, OBS, A = 4096, 16, 388, 80


class ActorCritic(nn.Module):                      # module.py:25-55 (shapes of cfg/ppo/config.yaml), act() of :73-87
    def __init__(self):
        super().__init__()
        self.asymmetric = False

        def mlp(out):
            return nn.Sequential(nn.Linear(OBS, 1024), nn.ELU(), nn.Linear(1024, 1024), nn.ELU(), nn.Linear(1024, 512), nn.ELU(), nn.Linear(512, out))
        self.actor, self.critic = mlp(A), mlp(1)
        self.log_std = nn.Parameter(torch.log(torch.tensor(0.8)) * torch.ones(A))

    @torch.no_grad()
    def act(self, obs, states):
        mean = self.actor(obs)
        cov = torch.diag(self.log_std.exp() * self.log_std.exp())
        dist = torch.distributions.MultivariateNormal(mean, scale_tril=cov)
        actions = dist.sample()
        return actions, dist.log_prob(actions), self.critic(obs), mean, self.log_std.repeat(mean.shape[0], 1)


torch.manual_seed(0)
ac = ActorCritic().to(dev)
fr = synthetic.ten_ant_frames(N, 32, seed=3)
out = {"envs": N, "horizon": T}
for label, env_cls, policy in (("eager", VecTaskPython, "fused"), ("graphed_env", GraphedVecTaskPython, "fused"),
                               ("graphed_rollout", VecTaskPython, "fused"), ("torch_fp32_policy", VecTaskPython, "torch")):
    task = TenAnt({"env": {"numEnvs": N, "env_name": "ten_ant"}, "sim": {"dt": 0.0166}, "seed": 1},
                  provider=ReplayProvider({"root": fr["root"], "dof": fr["dof"]}, device=dev))
    env = env_cls(task, dev)
    st = RolloutStorage(N, T, (OBS,), (0,), (A,), dev)
    pol = PPOActorCriticForward(ac, dev) if policy == "fused" else ac
    states = torch.zeros(N, 0, device=dev)
    current_obs = env.reset().clone()
    ro = GraphedPPORollout(env, pol, st, 0.99, 0.95) if label == "graphed_rollout" else None
    if ro is not None:
        ro.start_from(current_obs)

    def rollout():
        if ro is not None:
            ro.run()
            st.clear()
            return
        for _ in range(T):
            actions, logp, values, mu, sigma = pol.act(current_obs, states)
            next_obs, rews, dones, _ = env.step(actions)
            st.add_transitions(current_obs, states, actions, rews, dones, values, logp, mu, sigma)
            current_obs.copy_(next_obs)
        _, _, last_values, _, _ = pol.act(current_obs, states)
        st.compute_returns(last_values, 0.99, 0.95)
        st.clear()

    for _ in range(3):
        rollout()
    torch.cuda.synchronize()
    K = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(K):
        rollout()
    e1.record(); host = time.perf_counter() - t0
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    devt = e0.elapsed_time(e1) * 1e-3
    out[label] = {"host_us_per_env_step": host / (K * T) * 1e6, "device_us_per_env_step": devt / (K * T) * 1e6,
                  "env_steps_per_s": N * K * T / wall}
    print(label, out[label], flush=True)
    del env, task, st, pol, ro
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_ppo_rollout.json", "w"), indent=1)
