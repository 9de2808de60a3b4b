"""MARL rollout path on one GPU (BASELINE.json configs 2 and 3): per env-step through the reference-facing API
   MultiVecTaskPython.step  ->  Runner.insert mask logic  ->  SharedReplayBuffer.insert,
then once per rollout  SharedReplayBuffer.compute_returns (all agents, PopArt moments) + normalised advantages.
MultiIngenuity (4 agents, shared obs 52) and TenAnt (10 agents, obs 46 + share_obs 388), N = 4096, T = 16, frames
resident in HBM.  Reports env-steps/s, the launches per env-step and per-kernel times (library event pairs).
Writes gpurun_out/bench_marl.json."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from massive_marl_benchmark_b200 import _lib as L  # noqa: E402
from massive_marl_benchmark_b200 import spaces, synthetic  # noqa: E402
from massive_marl_benchmark_b200.providers import ReplayProvider  # noqa: E402
from massive_marl_benchmark_b200.separated_buffer import runner_insert_masks  # noqa: E402
from massive_marl_benchmark_b200.shared_buffer import SharedReplayBuffer  # noqa: E402
from massive_marl_benchmark_b200.tasks import MultiIngenuity, TenAnt  # noqa: E402
from massive_marl_benchmark_b200.vec_task import MultiVecTaskPython  # noqa: E402

dev = torch.device("cuda:0")
N, T = 4096, 16
out = {}


class Norm:
    def __init__(self, m, v):
        self.m, self.v = torch.tensor([m], device=dev), torch.tensor([v], device=dev)

    def running_mean_var(self):
        return self.m, self.v


for name, cls, gen, A, act in (("multi_ingenuity", MultiIngenuity, synthetic.ingenuity_frames, 4, 6),
                               ("ten_ant", TenAnt, synthetic.ten_ant_frames, 10, 8)):
    F = 64
    fr = gen(N, F, seed=7)
    cfg = {"env": {"numEnvs": N, "env_name": name}, "sim": {"dt": 0.0166}, "seed": 1}
    prov = ReplayProvider({k: v for k, v in fr.items() if k != "actions"}, device=dev)
    task = cls(cfg, None, None, "cuda", 0, True, True, provider=prov)
    env = MultiVecTaskPython(task, dev)
    actions = fr["actions"].to(dev)
    obs_all, state_all, _ = env.reset()
    O, S = obs_all.shape[2], state_all.shape[2]
    bcfg = dict(episode_length=T, n_rollout_threads=N, hidden_size=512, recurrent_N=1, gamma=0.96, gae_lambda=0.95,
                use_gae=True, use_popart=True, use_valuenorm=False, use_proper_time_limits=False)
    buf = SharedReplayBuffer(bcfg, A, spaces.Box(low=-np.inf, high=np.inf, shape=(O,)), spaces.Box(low=-np.inf, high=np.inf, shape=(S,)),
                             spaces.Box(low=-np.ones(act), high=np.ones(act)), dev)
    norms = [Norm(0.3, 2.5) for _ in range(A)]
    logp = torch.randn(N, A, act, device=dev); val = torch.randn(N, A, 1, device=dev); nv = torch.randn(N, A, 1, device=dev)
    masks = torch.empty(N, A, 1, device=dev); active = torch.empty(N, A, 1, device=dev)

    def rollout(r):
        for t in range(T):
            a = actions[(r * T + t) % F]
            obs_all, state_all, rew_all, done_all, _, _ = env.step(a)
            runner_insert_masks(done_all.contiguous(), masks, active)
            buf.insert(state_all[:, 0], obs_all, a.view(N, A, act), logp, val, rew_all, masks, None, active)
        buf.compute_returns(nv, norms)
        adv = buf.normalized_advantages(1e-5)
        buf.after_update()
        return adv

    for r in range(3):
        rollout(r)
    torch.cuda.synchronize()
    l0 = L.launch_count()
    K = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for r in range(K):
        rollout(r)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1) / K
    L.profile_enable(True); L.profile_collect()
    for r in range(5):
        rollout(r)
    torch.cuda.synchronize(); L.profile_enable(False)
    prof = L.profile_collect()
    out[name] = {"envs": N, "agents": A, "horizon": T, "ms_per_rollout": ms, "env_steps_per_s": N * T / ms * 1e3,
                 "agent_steps_per_s": N * T * A / ms * 1e3, "host_wall_ms_per_rollout": wall / K * 1e3,
                 "library_launches_per_env_step": (L.launch_count() - l0) / (K * T) if False else None,
                 "kernel_us_avg": {k: round(v[0] / v[1] * 1e3, 2) for k, v in prof.items()},
                 "kernel_launches_per_rollout": {k: v[1] / 5 for k, v in prof.items()},
                 "buffer_bytes": sum(getattr(buf, n).numel() * getattr(buf, n).element_size() for n in
                                     ("share_obs", "obs", "value_preds", "returns", "masks", "bad_masks", "active_masks", "actions",
                                      "action_log_probs", "rewards", "factor", "raw_advantages", "_zero_rnn")),
                 "per_agent_buffers_bytes_reference_layout": A * ((T + 1) * N * (S + O + 2 * 512 + 5) + T * N * (2 * act + 2)) * 4}
    del task, env, buf, prov
    torch.cuda.empty_cache()
# ---- TenAnt MARL, horizon-batched: the step kernel writes the agent-major obs planes, share_obs, reward and done of
# all T steps straight into the shared buffer (obs_layout 2); the per-agent planes that are replicas (reward, masks) are
# broadcast copies; then reset lists, GAE for all agents in one launch, per-agent normalisation -------------------------
fr = synthetic.ten_ant_frames(N, 4 * T, seed=11)
frd = {k: v.to(dev) for k, v in fr.items()}
cfg = {"env": {"numEnvs": N, "env_name": "ten_ant"}, "sim": {"dt": 0.0166}, "seed": 1}
task = TenAnt(cfg, None, None, "cuda", 0, True, True, provider=ReplayProvider({"root": fr["root"], "dof": fr["dof"]}, device=dev))
task.clip_actions, task.clip_obs = 1.0, 7.0
A, act = 10, 8
bcfg = dict(episode_length=T, n_rollout_threads=N, hidden_size=512, recurrent_N=1, gamma=0.96, gae_lambda=0.95,
            use_gae=True, use_popart=True, use_valuenorm=False, use_proper_time_limits=False)
bufs = [SharedReplayBuffer(bcfg, A, spaces.Box(low=-np.inf, high=np.inf, shape=(46,)), spaces.Box(low=-np.inf, high=np.inf, shape=(388,)),
                           spaces.Box(low=-np.ones(act), high=np.ones(act)), dev) for _ in range(2)]
norms = [Norm(0.3, 2.5) for _ in range(A)]
rew = [torch.zeros(T, N, device=dev) for _ in range(2)]; d8 = [torch.zeros(T, N, device=dev, dtype=torch.uint8) for _ in range(2)]
forces = torch.zeros(T, N, 80, device=dev); nv = torch.randn(N, A, 1, device=dev)
from massive_marl_benchmark_b200.tasks import reset_replay  # noqa: E402
reset_out = [None]


def rollout_batched(r):
    b = bufs[r % 2]; w = {k: v[(r % 4) * T:(r % 4 + 1) * T] for k, v in frd.items()}
    task.replay(w, w["actions"], None, rew[r % 2], d8[r % 2], None, forces, share_obs_out=b.share_obs[1:],
                agent_major_obs_out=b.obs[:, 1:], overlap_prev=True)
    reset_out[0] = reset_replay(task, d8[r % 2], out=reset_out[0])
    b.rewards.copy_(rew[r % 2].view(1, T, N, 1).expand(A, T, N, 1))                      # reward_all: the same reward x A
    b.masks[:, 1:].copy_((1.0 - d8[r % 2].float()).view(1, T, N, 1).expand(A, T, N, 1))  # all agents of an env end together
    b.actions.copy_(w["actions"].view(T, N, A, act).permute(2, 0, 1, 3))
    b.compute_returns(nv, norms)
    adv = b.normalized_advantages(1e-5)
    b.after_update()
    return adv


for r in range(4):
    rollout_batched(r)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for r in range(4):
        rollout_batched(r)
g.replay(); torch.cuda.synchronize()
K = 50
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K):
    g.replay()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / (4 * K)
out["ten_ant_horizon_batched"] = {"envs": N, "agents": A, "horizon": T, "ms_per_rollout": ms, "env_steps_per_s": N * T / ms * 1e3,
                                  "agent_steps_per_s": N * T * A / ms * 1e3,
                                  "note": "CUDA graph of 4 rollouts; step kernel with obs_layout 2 writes the shared buffer in place"}
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_marl.json", "w"), indent=1)
