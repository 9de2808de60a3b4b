"""Per-kernel throughput sweep (BASELINE.json configs 0/2/4): GAE + normalise and shuffle + gather at 1M..64M
transitions, OneAnt / MultiIngenuity / TenAnt step kernels at large N.  Reports algorithmic GB/s
against the measured HBM peak.  Writes gpurun_out/sweep_kernels.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from massive_marl_benchmark_b200 import _lib as L  # noqa: E402
from massive_marl_benchmark_b200 import synthetic  # noqa: E402
from massive_marl_benchmark_b200.providers import ReplayProvider  # noqa: E402
from massive_marl_benchmark_b200.storage import RolloutStorage  # noqa: E402
from massive_marl_benchmark_b200.tasks import MultiIngenuity, OneAnt, TenAnt  # noqa: E402

dev = torch.device("cuda:0")
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out = {"peak_gbs": PEAK, "gae": [], "gather": [], "tasks": []}
T = 16
ONLY = os.environ.get("SWEEP_ONLY", "")
for N in (() if ONLY == "tasks" else (65536, 262144, 1048576, 4194304)):
    st = RolloutStorage(N, T, (60,), (0,), (8,), dev, "random")
    st.rewards.normal_(); st.values.normal_(); st.dones.copy_((torch.rand(T, N, 1, device=dev) < 0.01).to(torch.uint8))
    lv = torch.randn(N, 1, device=dev)
    ms = timeit(lambda: st.compute_returns(lv, 0.96, 0.95))
    nbytes = 25 * T * N
    out["gae"].append({"transitions": T * N, "ms": ms, "gbs": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / PEAK,
                       "transitions_per_s": T * N / ms * 1e3})
    # one epoch of shuffle + gather of all fields, 4 minibatches (payload 60*4 + 3*32 + 20 = 356 B per row)
    mb = T * N // 4
    bufs = None
    def epoch():
        global bufs
        for idx in st.mini_batch_generator(4):
            bufs = st.gather_minibatch(idx, bufs)
    ms = timeit(epoch, iters=5, warm=2)
    row = 60 * 4 + 8 * 4 * 3 + 5 * 4
    nbytes = (8 + 2 * row) * T * N + 8 * T * N          # int64 index read + payload read+write (+ permutation write)
    out["gather"].append({"transitions": T * N, "ms_per_epoch": ms, "gbs": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / PEAK})
    del st, bufs
    bufs = None
    torch.cuda.empty_cache()

for name, cls, gen, nbytes_step, N, Tt in (("ten_ant", TenAnt, synthetic.ten_ant_frames, 3409, 16384, 8),
                                          ("one_ant", OneAnt, synthetic.one_ant_frames, 672, 262144, 8),
                                          ("multi_ingenuity", MultiIngenuity, synthetic.ingenuity_frames, 836, 262144, 8)):
    fr = gen(N, Tt, seed=1)
    frd = {k: v.to(dev) for k, v in fr.items()}
    cfg = {"env": {"numEnvs": N, "env_name": name}, "sim": {"dt": 0.0166}, "seed": 1}
    prov = ReplayProvider({k: v for k, v in fr.items() if k != "actions"}, device=dev)
    task = cls(cfg, provider=prov)
    task.clip_actions, task.clip_obs = 1.0, 5.0
    W = task.num_obs
    obs = torch.zeros(Tt, N, W, device=dev); rew = torch.zeros(Tt, N, device=dev)
    d8 = torch.zeros(Tt, N, device=dev, dtype=torch.uint8)
    forces = torch.zeros(Tt, N, 80 if name == "ten_ant" else (8 if name == "one_ant" else 72), device=dev)
    if name == "multi_ingenuity":
        forces = forces.view(Tt, N, 24, 3)
    ms = timeit(lambda: task.replay(frd, frd["actions"], obs, rew, d8, None, forces), iters=10)
    nb = nbytes_step * Tt * N
    out["tasks"].append({"task": name, "envs": N, "frames": Tt, "ms": ms, "env_steps_per_s": Tt * N / ms * 1e3,
                         "gbs": nb / ms / 1e6, "frac": nb / ms / 1e6 / PEAK})
    del task, obs, forces, frd
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep_kernels.json", "w"), indent=1)
