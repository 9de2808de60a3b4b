"""Small invocation of every kernel family, sized for compute-sanitizer (memcheck / racecheck / synccheck) on a
workstation GPU.  NOT for the shared B200 pool: compute-sanitizer is closed there (runs under it have left GPUs needing a
reset); on the pool this script only serves as a plain smoke run of all kernel families."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from massive_marl_benchmark_b200 import synthetic
from massive_marl_benchmark_b200.providers import ReplayProvider
from massive_marl_benchmark_b200.storage import RolloutStorage
from massive_marl_benchmark_b200.tasks import MultiIngenuity, OneAnt, TenAnt, reset_replay
from massive_marl_benchmark_b200.episodes import EpisodeTracker
from massive_marl_benchmark_b200.mlp import FusedMLP
dev = torch.device("cuda:0")
N, T = 100, 5
for name, cls, gen in (("ten_ant", TenAnt, synthetic.ten_ant_frames), ("one_ant", OneAnt, synthetic.one_ant_frames),
                       ("multi_ingenuity", MultiIngenuity, synthetic.ingenuity_frames)):
    fr = gen(N, T, seed=3)
    frd = {k: v.to(dev) for k, v in fr.items()}
    cfg = {"env": {"numEnvs": N, "env_name": name}, "sim": {"dt": 0.0166}, "seed": 1}
    task = cls(cfg, provider=ReplayProvider({k: v for k, v in fr.items() if k != "actions"}, device=dev))
    task.clip_actions, task.clip_obs = 1.0, 5.0
    W = task.num_obs
    obs = torch.zeros(T, N, W, device=dev); rew = torch.zeros(T, N, device=dev); d8 = torch.zeros(T, N, device=dev, dtype=torch.uint8)
    fo = torch.zeros(T, N, 80 if name == "ten_ant" else (8 if name == "one_ant" else 72), device=dev)
    if name == "multi_ingenuity": fo = fo.view(T, N, 24, 3)
    for rep in range(2):
        if name == "ten_ant":
            task.replay(frd, frd["actions"], obs, rew, d8, None, fo, overlap_prev=True)
        else:
            task.replay(frd, frd["actions"], obs, rew, d8, None, fo)
    reset_replay(task, d8)
    task.step(frd["actions"][0])
    torch.cuda.synchronize()
    print(name, "ok", float(rew.sum()))
st = RolloutStorage(N, T, (388,), (0,), (80,), dev, "random")
st.rewards.normal_(); st.values.normal_()
st.compute_returns(torch.randn(N, 1, device=dev), 0.96, 0.95)
for idx in st.mini_batch_generator(2):
    st.gather_minibatch(idx)
big = RolloutStorage(8192, 4, (4,), (0,), (2,), dev)
big.rewards.normal_(); big.values.normal_(); big.compute_returns(torch.randn(8192, 1, device=dev), 0.96, 0.95)
tr = EpisodeTracker(N, dev); tr.update(st.rewards, (torch.rand(T, N, 1, device=dev) < 0.3).to(torch.uint8)); tr.means()
net = torch.nn.Sequential(torch.nn.Linear(60, 256), torch.nn.ELU(), torch.nn.Linear(256, 128), torch.nn.ELU(), torch.nn.Linear(128, 8)).to(dev)
y = FusedMLP.from_sequential(net, dev)(torch.randn(300, 60, device=dev))
torch.cuda.synchronize()
print("storage / episodes / mlp ok", float(y.abs().sum()))
