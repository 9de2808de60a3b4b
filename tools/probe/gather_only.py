"""Shuffle + gather epochs alone at the sweep sizes (config 5 payload: 60-float obs, 3 x 8-float, 5 scalars)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from massive_marl_benchmark_b200.storage import RolloutStorage
dev = torch.device("cuda:0")
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PEAK = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) else 6540.0
T = 16
for N in (65536, 262144, 1048576):
    st = RolloutStorage(N, T, (60,), (0,), (8,), dev, "random")
    bufs = [None]
    def epoch():
        for idx in st.mini_batch_generator(4):
            bufs[0] = st.gather_minibatch(idx, bufs[0])
    for _ in range(2): epoch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): epoch()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    row = 60 * 4 + 8 * 4 * 3 + 5 * 4
    nb = (8 + 2 * row) * T * N + 8 * T * N
    print("transitions %9d  epoch %.3f ms  %.0f GB/s algorithmic (%.2f of peak)" % (T * N, ms, nb / ms / 1e6, nb / ms / 1e6 / PEAK), flush=True)
    del st; bufs[0] = None; torch.cuda.empty_cache()
