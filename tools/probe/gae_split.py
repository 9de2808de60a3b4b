"""GAE scan and advantage normalisation timed separately at large sizes (algorithmic GB/s vs measured peak)."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from massive_marl_benchmark_b200.storage import RolloutStorage
dev = torch.device("cuda:0")
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PEAK = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) else 6540.0
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
T = 16
for N in (262144, 1048576, 4194304):
    st = RolloutStorage(N, T, (1,), (0,), (1,), dev)
    st.rewards.normal_(); st.values.normal_(); st.dones.copy_((torch.rand(T, N, 1, device=dev) < 0.01).to(torch.uint8))
    lv = torch.randn(N, 1, device=dev)
    a = timeit(lambda: st.compute_returns_scan(lv, 0.96, 0.95))
    b = timeit(lambda: st.normalize_advantages())
    print("N=%8d  scan %.3f ms %.0f GB/s (%.2f)   normalise %.3f ms %.0f GB/s (%.2f)" % (
        N, a, 17 * T * N / a / 1e6, 17 * T * N / a / 1e6 / PEAK, b, 8 * T * N / b / 1e6, 8 * T * N / b / 1e6 / PEAK), flush=True)
    del st
    torch.cuda.empty_cache()
