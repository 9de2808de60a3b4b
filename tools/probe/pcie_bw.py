"""Host->device bandwidth of back-to-back pinned copies of one packed TenAnt frame (6.27 MB) - the ceiling of the e2e loop."""
import torch
dev = torch.device("cuda:0")
n = (45056 * 13 + 327680 * 2 + 4096 * 80)
h = torch.empty(16, n, dtype=torch.float32).pin_memory()
d = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(2)]
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for i in range(8): d[i & 1].copy_(h[i % 16], non_blocking=True)
    s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    R = 200
    for i in range(R): d[i & 1].copy_(h[i % 16], non_blocking=True)
    e1.record(s)
    s.synchronize()
ms = e0.elapsed_time(e1) / R
print("frame %.2f MB: %.1f us per copy = %.1f GB/s -> ceiling %.1f M env-steps/s at 4096 envs" % (n * 4 / 1e6, ms * 1e3, n * 4 / ms / 1e6, 4096 / ms / 1e3))
big = torch.empty(64 * 1024 * 1024, dtype=torch.float32).pin_memory(); dbig = torch.empty_like(big, device=dev)
torch.cuda.synchronize(); e0.record(); dbig.copy_(big, non_blocking=True); e1.record(); torch.cuda.synchronize()
print("256 MB copy: %.1f GB/s" % (big.numel() * 4 / e0.elapsed_time(e1) / 1e6))
