"""Launch list of one GraphedPPORollout (TenAnt N = 4096, T = 16) for `ncu --metrics gpu__time_duration.sum`: one eager rollout,
two captures, two replays.  The last ~86 launches of the list are one replayed rollout."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from massive_marl_benchmark_b200 import synthetic  # noqa: E402
from massive_marl_benchmark_b200.mlp import PPOActorCriticForward  # noqa: E402
from massive_marl_benchmark_b200.ppo_rollout import GraphedPPORollout  # noqa: E402
from massive_marl_benchmark_b200.providers import ReplayProvider  # noqa: E402
from massive_marl_benchmark_b200.storage import RolloutStorage  # noqa: E402
from massive_marl_benchmark_b200.tasks import TenAnt  # noqa: E402
from massive_marl_benchmark_b200.vec_task import VecTaskPython  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

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
fr = synthetic.ten_ant_frames(N, 32, seed=3)
task = TenAnt({"env": {"numEnvs": N, "env_name": "ten_ant"}, "sim": {"dt": 0.0166}, "seed": 1},
              provider=ReplayProvider({"root": fr["root"], "dof": fr["dof"]}, device=dev))
env = VecTaskPython(task, dev)
st = RolloutStorage(N, T, (388,), (0,), (80,), dev)
ro = GraphedPPORollout(env, PPOActorCriticForward(ac, dev), st, 0.99, 0.95)
for _ in range(int(os.environ.get("ROLLOUTS", "5"))):
    ro.run()
    st.clear()
torch.cuda.synchronize()
print("captures", ro.captures)
