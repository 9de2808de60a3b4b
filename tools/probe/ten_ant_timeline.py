"""Per-CTA phase timeline of the TenAnt step kernel.  Needs a trace build of the library:
    make -C massive_marl_benchmark_b200/csrc clean && make -C massive_marl_benchmark_b200/csrc EXTRA=-DMMB_TRACE
(rebuild without EXTRA afterwards: the trace symbol is not part of the ABI)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from massive_marl_benchmark_b200 import synthetic
from massive_marl_benchmark_b200.providers import ReplayProvider
from massive_marl_benchmark_b200.tasks import TenAnt
dev = torch.device("cuda:0"); N, T, sets = 4096, 16, 4
frs = [{k: v.to(dev) for k, v in synthetic.ten_ant_frames(N, T, seed=s).items()} for s in range(sets)]
cfg = {"env": {"numEnvs": N, "env_name": "ten_ant"}, "sim": {"dt": 0.0166}, "seed": 1}
task = TenAnt(cfg, provider=ReplayProvider({k: v.cpu() for k, v in frs[0].items() if k != "actions"}, device=dev))
task.clip_actions, task.clip_obs = 1.0, 5.0
outs = [(torch.zeros(T, N, 388, device=dev), torch.zeros(T, N, device=dev), torch.zeros(T, N, device=dev, dtype=torch.uint8),
         torch.zeros(T, N, 80, device=dev)) for _ in range(sets)]
for i in range(12):
    f, o = frs[i % sets], outs[i % sets]
    task.replay(f, f["actions"], o[0], o[1], o[2], None, o[3], overlap_prev=True)
torch.cuda.synchronize()
lib = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "massive_marl_benchmark_b200", "libmmb_b200.so"))
buf = (C.c_ulonglong * (64 * 16))()
lib.mmb_dbg_trace(buf)
names = ["start", "dof:B1", "dof:own work done", "dof:B2 arrive", "dof:B2 passed", "core:root ready", "core:B2 arrive", "core:postB2 done",
         "dof:postB2 done", "B3 passed", "finish done", "exit"]
for c in range(4):
    t = list(buf)[c * 16:c * 16 + 16]
    print("   finish: enter %d, after griddep_wait %d, sums done %d, reward stored %d" % (t[14] - t[0], t[15] - t[0], t[12] - t[0], t[13] - t[0]))
    t = t[:12]
    print("tile %3d frame 5 (ns): " % (c * 64) + ", ".join("%s %d" % (n, v - t[0]) for n, v in zip(names, t)))
