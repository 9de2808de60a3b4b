"""MARL rollout collection through `runner.Runner` (BASELINE configs[2] / [3] per-step path): one env step =
`collect` (all agents' actor + critic forwards, sampling) -> `MultiVecTaskPython.step` -> `insert`, then `compute` once per
episode.  Compares the reference's structure (2 x num_agents torch module forwards per step, per-agent buffers) with
`team_forward=True, shared_buffer=True` (two grouped tcgen05 forwards, one shared insert).  TenAnt MAPPO and MultiIngenuity
MAPPO, N = 4096.  Needs baseline/_ref (the reference's policies / trainers).  Writes gpurun_out/bench_runner.json."""
import contextlib
import io
import json
import os
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "baseline", "_ref")
from oracle import refshim  # noqa: E402
refshim.install(REF)
import yaml  # noqa: E402
from massive_marl_benchmark_b200 import synthetic  # noqa: E402
from massive_marl_benchmark_b200.providers import ReplayProvider  # noqa: E402
from massive_marl_benchmark_b200.runner import Runner  # noqa: E402
from massive_marl_benchmark_b200.tasks import MultiIngenuity, TenAnt  # noqa: E402
from massive_marl_benchmark_b200.vec_task import MultiVecTaskPython  # noqa: E402

dev = torch.device("cuda", 0)
N = int(os.environ.get("RUNNER_ENVS", "4096"))
out = {"envs": N}
for name, cls, gen, A in (("ten_ant", TenAnt, synthetic.ten_ant_frames, 10), ("multi_ingenuity", MultiIngenuity, synthetic.ingenuity_frames, 4)):
    config = yaml.safe_load(open(os.path.join(REF, "cfg", "mappo", "config.yaml")))
    T = config["episode_length"]
    fr = gen(N, 64, seed=5)
    res = {}
    for label, kw in (("reference_structure", dict()), ("team_forward_shared_buffer", dict(team_forward=True, shared_buffer=True))):
        cfg = {"env": {"numEnvs": N, "env_name": name}, "sim": {"dt": 0.0166}, "seed": 1}
        task = cls(cfg, None, None, "cuda", 0, True, True, provider=ReplayProvider({k: v for k, v in fr.items() if k != "actions"}, device=dev))
        env = MultiVecTaskPython(task, "cuda:0")
        config.update(n_rollout_threads=N, n_eval_rollout_threads=N, run_dir=tempfile.mkdtemp(), experiment_name="bench", use_eval=False)
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            r = Runner(vec_env=env, config=dict(config), model_dir="", writer=False, **kw)
        r.warmup()

        def episode():
            for step in range(T):
                values, actions, logps, rnn, rnn_c = r.collect(step)
                obs, share_obs, rewards, dones, infos, _ = env.step(actions)
                r.insert((obs, share_obs, rewards, dones, infos, values, actions, logps, rnn, rnn_c))
            r.compute()
            for b in (r.buffer if r.shared is None else []):
                b.after_update()
            if r.shared is not None:
                r.shared.after_update()

        for _ in range(3):
            episode()
        torch.cuda.synchronize()
        K = 10
        t0 = time.perf_counter()
        for _ in range(K):
            episode()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / (K * T)
        res[label] = {"us_per_env_step": dt * 1e6, "env_steps_per_s": N / dt, "agent_steps_per_s": N * A / dt}
        del r, env, task
        torch.cuda.empty_cache()
    res["speedup"] = res["reference_structure"]["us_per_env_step"] / res["team_forward_shared_buffer"]["us_per_env_step"]
    out[name] = res
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_runner.json"), "w"), indent=1)
