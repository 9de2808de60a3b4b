"""Where the HOST time of one MARL Runner env step goes (TenAnt, team_forward + shared_buffer, N = 4096): cProfile over 40 steps,
plus the GPU-side time of the same steps (CUDA events), so host-bound vs device-bound is visible."""
import contextlib, cProfile, io, os, pstats, sys, tempfile, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "baseline", "_ref")
from oracle import refshim  # noqa: E402
refshim.install(REF)
import yaml  # noqa: E402
from massive_marl_benchmark_b200 import synthetic  # noqa: E402
from massive_marl_benchmark_b200.providers import ReplayProvider  # noqa: E402
from massive_marl_benchmark_b200.runner import Runner  # noqa: E402
from massive_marl_benchmark_b200.tasks import TenAnt  # noqa: E402
from massive_marl_benchmark_b200.vec_task import MultiVecTaskPython  # noqa: E402

dev = torch.device("cuda", 0)
N = 4096
config = yaml.safe_load(open(os.path.join(REF, "cfg", "mappo", "config.yaml")))
fr = synthetic.ten_ant_frames(N, 64, seed=5)
cfg = {"env": {"numEnvs": N, "env_name": "ten_ant"}, "sim": {"dt": 0.0166}, "seed": 1}
task = TenAnt(cfg, None, None, "cuda", 0, True, True, provider=ReplayProvider({k: v for k, v in fr.items() if k != "actions"}, device=dev))
env = MultiVecTaskPython(task, "cuda:0")
config.update(n_rollout_threads=N, n_eval_rollout_threads=N, run_dir=tempfile.mkdtemp(), experiment_name="bench", use_eval=False)
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    r = Runner(vec_env=env, config=dict(config), model_dir="", writer=False, team_forward=True, shared_buffer=True)
r.warmup()


def steps(n, timers=None):
    for step in range(n):
        t = [time.perf_counter()]
        values, actions, logps, rnn, rnn_c = r.collect(step)
        t.append(time.perf_counter())
        obs, share_obs, rewards, dones, infos, _ = env.step(actions)
        t.append(time.perf_counter())
        r.insert((obs, share_obs, rewards, dones, infos, values, actions, logps, rnn, rnn_c))
        t.append(time.perf_counter())
        if timers is not None:
            for k in range(3):
                timers[k] += t[k + 1] - t[k]


steps(8)
torch.cuda.synchronize()
tm = [0.0, 0.0, 0.0]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
steps(7, tm)
e1.record(); host = time.perf_counter() - t0
torch.cuda.synchronize()
print("per step: host enqueue %.0f us (collect %.0f, env.step %.0f, insert %.0f), device %.0f us" %
      (host / 7 * 1e6, tm[0] / 7 * 1e6, tm[1] / 7 * 1e6, tm[2] / 7 * 1e6, e0.elapsed_time(e1) / 7 * 1e3))
# each part alone, 20 calls back to back: host time per call against device time per call (device < host: host-bound)
values, actions, logps, rnn, rnn_c = r.collect(0)
obs, share_obs, rewards, dones, infos, _ = env.step(actions)
parts = {"collect": lambda: r.collect(1), "team.get_actions": None, "env.step": lambda: env.step(actions)}
import massive_marl_benchmark_b200.runner as _rn
for name, fn in parts.items():
    if fn is None:
        continue
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e0.record()
    for _ in range(20):
        fn()
    e1.record(); h = time.perf_counter() - t0
    torch.cuda.synchronize()
    print("%s: host %.0f us, device %.0f us per call" % (name, h / 20 * 1e6, e0.elapsed_time(e1) / 20 * 1e3))
if getattr(r, "team", None) is not None:
    so, ob = r.shared.share_obs[0], r.shared.obs[:, 0] if hasattr(r.shared, "obs") else None
r.shared.after_update() if r.shared is not None else None
pr = cProfile.Profile()
pr.enable()
steps(7)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print("\n".join(l[:150] for l in s.getvalue().splitlines()[:50]))
