"""TenAnt step kernel: time per launch vs problem size (fixed launch overhead vs asymptotic per-unit cost)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from massive_marl_benchmark_b200 import synthetic
from massive_marl_benchmark_b200.providers import ReplayProvider
from massive_marl_benchmark_b200.tasks import TenAnt

dev = torch.device("cuda:0")
T = 16
for N in (2048, 4096, 8192, 16384, 32768):
    sets = 2 if N >= 16384 else 4
    frs = []
    for s in range(sets):
        fr = synthetic.ten_ant_frames(N, T, seed=s)
        frs.append({k: v.to(dev) for k, v in fr.items()})
    cfg = {"env": {"numEnvs": N, "env_name": "ten_ant"}, "sim": {"dt": 0.0166}, "seed": 1}
    task = TenAnt(cfg, provider=ReplayProvider({k: v for k, v in fr.items() if k != "actions"}, device=dev))
    task.clip_actions, task.clip_obs = 1.0, 5.0
    outs = [(torch.zeros(T, N, 388, device=dev), torch.zeros(T, N, device=dev), torch.zeros(T, N, device=dev, dtype=torch.uint8),
             torch.zeros(T, N, 80, device=dev)) for _ in range(sets)]
    def run(i):
        f = frs[i % sets]; o = outs[i % sets]
        task.replay(f, f["actions"], o[0], o[1], o[2], None, o[3])
    for i in range(sets): run(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(sets): run(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 40
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * sets)
    nb = 3409 * T * N
    print("N=%6d  %.2f us/launch  %.3f ns/env-step  %.0f GB/s" % (N, us, us * 1e3 / (T * N), nb / us / 1e3), flush=True)
    del task, outs, frs
    torch.cuda.empty_cache()
