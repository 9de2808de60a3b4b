"""Where the device time of a graphed PPO rollout step goes (TenAnt N = 4096, T = 16): graphs of 16 x [a subset of the step's
launches], replayed, CUDA events.  act = dual-network chain + sampling; env = step kernel (reset compaction on the side branch);
ins = insert on the side branch."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from massive_marl_benchmark_b200 import synthetic  # noqa: E402
from massive_marl_benchmark_b200.mlp import PPOActorCriticForward  # noqa: E402
from massive_marl_benchmark_b200.providers import ReplayProvider  # noqa: E402
from massive_marl_benchmark_b200.storage import RolloutStorage  # noqa: E402
from massive_marl_benchmark_b200.tasks import TenAnt  # noqa: E402
from massive_marl_benchmark_b200.vec_task import VecTaskPython  # noqa: E402

dev = torch.device("cuda:0")
N, T = 4096, 16


def net(out):
    import torch.nn as nn
    return nn.Sequential(nn.Linear(388, 1024), nn.ELU(), nn.Linear(1024, 1024), nn.ELU(), nn.Linear(1024, 512), nn.ELU(), nn.Linear(512, out))


class AC(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.asymmetric = False
        self.actor, self.critic = net(80), net(1)
        self.log_std = torch.nn.Parameter(torch.log(torch.tensor(0.8)) * torch.ones(80))


torch.manual_seed(0)
ac = AC().to(dev)
fr = synthetic.ten_ant_frames(N, 16, seed=3)
task = TenAnt({"env": {"numEnvs": N, "env_name": "ten_ant"}, "sim": {"dt": 0.0166}, "seed": 1},
              provider=ReplayProvider({"root": fr["root"], "dof": fr["dof"]}, device=dev))
env = VecTaskPython(task, dev)
st = RolloutStorage(N, T, (388,), (0,), (80,), dev)
pol = PPOActorCriticForward(ac, dev)
pol.use_device_step_counter()
task.use_device_step_counter()
states = torch.zeros(N, 0, device=dev)
obs = env.reset().clone()
obs2 = torch.empty_like(obs)
side = torch.cuda.Stream()
fixed_actions = torch.zeros(N, 80, device=dev)
bufs = [obs, obs2]


def body(parts):
    main = torch.cuda.current_stream()
    flip, keep, joined = 0, [], None

    def out(*shape):
        nonlocal flip
        flip ^= 1
        return bufs[flip]
    task._fresh_out = out
    try:
        cur = bufs[0]
        for _ in range(T):
            if "act" in parts:
                o = pol.act(cur, states)
                keep.append(o)
                actions, logp, values, mu, sigma = o
            elif "mlp" in parts:
                o = pol._mean_value(cur, states)
                keep.append(o)
                actions = fixed_actions
            else:
                actions = fixed_actions
            if "env" in parts:
                if joined is not None:
                    main.wait_event(joined)
                task.reset_ahead = "ahead" in parts
                task.step(actions)
                task.reset_ahead = False
                ev = torch.cuda.Event(); ev.record(main); side.wait_event(ev)
                with torch.cuda.stream(side):
                    if "ahead" in parts:
                        task.reset_idx()
                    if "ins" in parts and "act" in parts:
                        st.add_transitions(cur, states, actions, task.rew_buf, task.reset_buf, values, logp, mu, sigma)
                        st.step = 0 if st.step >= T else st.step
                    joined = torch.cuda.Event(); joined.record(side)
                cur = task.obs_clamped
        if joined is not None:
            main.wait_event(joined)
    finally:
        del task._fresh_out


res = {}
for name, parts in (("mlp", {"mlp"}), ("act", {"act"}), ("env_inline_reset", {"env"}), ("env_reset_ahead", {"env", "ahead"}),
                    ("act+env", {"act", "env", "ahead"}), ("act+env+ins", {"act", "env", "ahead", "ins"})):
    body(parts)                      # eager once
    st.step = 0
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body(parts)
    st.step = 0
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    res[name] = e0.elapsed_time(e1) * 1e3 / (20 * T)
    print(name, "%.2f us per step" % res[name], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/rollout_parts.json", "w"), indent=1)
