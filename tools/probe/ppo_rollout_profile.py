import cProfile, pstats, io, os, sys, time, torch, torch.nn as nn
sys.path.insert(0, "/root/repo")
from massive_marl_benchmark_b200 import synthetic
from massive_marl_benchmark_b200.mlp import PPOActorCriticForward
from massive_marl_benchmark_b200.providers import ReplayProvider
from massive_marl_benchmark_b200.storage import RolloutStorage
from massive_marl_benchmark_b200.tasks import TenAnt
from massive_marl_benchmark_b200.vec_task import GraphedVecTaskPython
dev = torch.device("cuda:0"); N, T, OBS, A = 4096, 16, 388, 80
class AC(nn.Module):
    def __init__(self):
        super().__init__(); self.asymmetric = False
        def mlp(o): return nn.Sequential(nn.Linear(OBS,1024), nn.ELU(), nn.Linear(1024,1024), nn.ELU(), nn.Linear(1024,512), nn.ELU(), nn.Linear(512,o))
        self.actor, self.critic = mlp(A), mlp(1); self.log_std = nn.Parameter(torch.zeros(A))
ac = AC().to(dev); fr = synthetic.ten_ant_frames(N, 32, seed=3)
task = TenAnt({"env": {"numEnvs": N, "env_name": "ten_ant"}, "sim": {"dt": 0.0166}, "seed": 1}, provider=ReplayProvider({"root": fr["root"], "dof": fr["dof"]}, device=dev))
env = GraphedVecTaskPython(task, dev); st = RolloutStorage(N, T, (OBS,), (0,), (A,), dev); pol = PPOActorCriticForward(ac, dev)
states = torch.zeros(N, 0, device=dev); cur = env.reset().clone()
tm = [0.0]*4
def rollout(timed=False):
    for _ in range(T):
        t0=time.perf_counter(); actions, logp, values, mu, sigma = pol.act(cur, states)
        t1=time.perf_counter(); nobs, rews, dones, _ = env.step(actions)
        t2=time.perf_counter(); st.add_transitions(cur, states, actions, rews, dones, values, logp, mu, sigma)
        t3=time.perf_counter(); cur.copy_(nobs); t4=time.perf_counter()
        if timed:
            for k,(a,b) in enumerate(((t0,t1),(t1,t2),(t2,t3),(t3,t4))): tm[k]+=b-a
    st.clear()
for _ in range(3): rollout()
torch.cuda.synchronize()
for _ in range(5): rollout(True)
torch.cuda.synchronize()
print("per step host us: act %.1f env.step %.1f add %.1f copy %.1f" % tuple(x/(5*T)*1e6 for x in tm))
pr = cProfile.Profile(); pr.enable(); rollout(); pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
print("\n".join(l[:140] for l in s.getvalue().splitlines()[:40]))
