"""Repeats the split-vs-mono comparison of tests/test_gpu_ten_ant.py::test_ten_ant_long_horizon_fallback_and_kernel_variants and
prints WHERE the outputs differ when they do (variant runs in subprocesses: MMB_TEN_ANT_VARIANT is read once per process)."""
import os
import subprocess
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
script = r'''
import sys, torch
sys.path.insert(0, %r)
from massive_marl_benchmark_b200 import synthetic
from massive_marl_benchmark_b200.providers import ReplayProvider
from massive_marl_benchmark_b200.tasks import TenAnt
dev = torch.device("cuda", 0)
N, T = 523, 12
fr = synthetic.ten_ant_frames(N, T, seed=77, fall_prob=0.02)
frd = {k: v.to(dev) for k, v in fr.items()}
outs = []
for rep in range(int(sys.argv[2])):
    task = TenAnt({"env": {"numEnvs": N, "env_name": "ten_ant"}, "sim": {"dt": 0.0166}, "seed": 1}, None, None, "cuda", 0, True, False,
                  provider=ReplayProvider({"root": fr["root"], "dof": fr["dof"]}, device=dev))
    task.clip_actions, task.clip_obs = 1.0, 5.0
    obs = torch.zeros(T, N, 388, device=dev); rew = torch.zeros(T, N, device=dev); d8 = torch.zeros(T, N, device=dev, dtype=torch.uint8)
    fo = torch.zeros(T, N, 80, device=dev)
    per = []
    for _ in range(2):
        task.replay(frd, frd["actions"], obs, rew, d8, None, fo)
        per.append(obs.cpu().clone())
    outs.append({"obs1": per[0], "obs": per[1], "rew": rew.cpu(), "d8": d8.cpu(), "fo": fo.cpu(), "prog": task.progress_buf.cpu(),
                 "pos": task.pos_before.cpu(), "goal": task.goal_before.cpu()})
torch.save(outs, sys.argv[1])
''' % (ROOT,)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
res = {}
with tempfile.TemporaryDirectory() as tmp:
    for variant in ("split", "mono"):
        out = os.path.join(tmp, variant + ".pt")
        subprocess.run([sys.executable, "-c", script, out, str(reps)], check=True, env=dict(os.environ, MMB_TEN_ANT_VARIANT=variant), timeout=600)
        res[variant] = torch.load(out)
bad = 0
ref = res["mono"][0]
for variant in ("split", "mono"):
    for i, r in enumerate(res[variant]):
        for k in r:
            if not torch.equal(r[k], ref[k]):
                bad += 1
                d = (r[k] != ref[k])
                idx = d.nonzero()
                print(variant, "rep", i, k, "differs in", int(d.sum()), "elements; first:", idx[:6].tolist(),
                      "values", r[k][d][:6].tolist(), "vs", ref[k][d][:6].tolist(), flush=True)
print("mismatches:", bad, "of", 2 * reps, "runs x", len(ref), "tensors")
