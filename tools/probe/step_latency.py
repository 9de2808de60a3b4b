"""Interactive per-step API (VecTaskPython.step = reset compaction + fused step kernel, frames resident in HBM): time per
env step and env-steps/s vs N, with the launch-bound small-N end reported as such (hard part 1b/1c).
Also the horizon-batched kernel over the same N (un-chained single launches of T = 16 frames)."""
import json, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from massive_marl_benchmark_b200 import synthetic
from massive_marl_benchmark_b200.providers import ReplayProvider
from massive_marl_benchmark_b200.tasks import TenAnt
from massive_marl_benchmark_b200.vec_task import GraphedVecTaskPython, VecTaskPython
dev = torch.device("cuda:0")
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PEAK = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) else 6540.0
out = []
for N in (4096, 16384, 65536, 262144):
    F = 16
    fr = synthetic.ten_ant_frames(N, F, seed=1)
    frd = {k: v.to(dev) for k, v in fr.items()}
    cfg = {"env": {"numEnvs": N, "env_name": "ten_ant"}, "sim": {"dt": 0.0166}, "seed": 1}
    task = TenAnt(cfg, provider=ReplayProvider({"root": fr["root"], "dof": fr["dof"]}, device=dev))
    env = VecTaskPython(task, dev)
    acts = frd["actions"]
    for i in range(20): env.step(acts[i % F])
    torch.cuda.synchronize()
    K = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(K): env.step(acts[i % F])
    e1.record(); torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / K * 1e6
    us = e0.elapsed_time(e1) / K * 1e3
    # the CUDA-graphed step (one graph per ring slot; actions written into the static tensor by the caller / passed in)
    tg = TenAnt({"env": {"numEnvs": N, "env_name": "ten_ant"}, "sim": {"dt": 0.0166}, "seed": 1},
                provider=ReplayProvider({"root": fr["root"], "dof": fr["dof"]}, device=dev))
    genv = GraphedVecTaskPython(tg, dev)
    for i in range(2 * F + 4): genv.step(acts[i % F])
    torch.cuda.synchronize()
    res = {}
    for name, arg in (("copy", lambda i: acts[i % F]), ("static", lambda i: None)):
        t0 = time.perf_counter(); e0.record()
        for i in range(K): genv.step(arg(i))
        e1.record(); torch.cuda.synchronize()
        res[name] = (e0.elapsed_time(e1) / K * 1e3, (time.perf_counter() - t0) / K * 1e6)
    del tg, genv
    T = 16
    obs = torch.zeros(T, N, 388, device=dev); rew = torch.zeros(T, N, device=dev); d8 = torch.zeros(T, N, device=dev, dtype=torch.uint8)
    fo = torch.zeros(T, N, 80, device=dev)
    for _ in range(3): task.replay(frd, acts, obs, rew, d8, None, fo)
    torch.cuda.synchronize(); e0.record()
    R = 20
    for _ in range(R): task.replay(frd, acts, obs, rew, d8, None, fo)
    e1.record(); torch.cuda.synchronize()
    ub = e0.elapsed_time(e1) / R * 1e3
    row = {"envs": N, "graphed_step_us": round(res["static"][0], 1), "graphed_step_host_us": round(res["static"][1], 1),
           "graphed_step_with_action_copy_us": round(res["copy"][0], 1),
           "graphed_step_env_steps_per_s": N / res["static"][0] * 1e6, "graphed_step_frac_of_hbm_peak": 3456 * N / res["static"][0] / 1e3 / PEAK,
           "per_step_api_us": round(us, 1), "per_step_api_host_us": round(wall, 1), "per_step_api_env_steps_per_s": N / us * 1e6,
           "per_step_api_frac_of_hbm_peak": 3456 * N / us / 1e3 / PEAK, "batched16_us_per_launch": round(ub, 1),
           "batched16_env_steps_per_s": T * N / ub * 1e6, "batched16_frac_of_hbm_peak": 3409 * T * N / ub / 1e3 / PEAK}
    out.append(row); print(row, flush=True)
    del task, env, obs, fo, frd; torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/step_latency.json", "w"), indent=1)
