"""PPO actor MLP forward (388-1024-1024-512-80) vs batch size: tcgen05 path, torch fp32 (the reference's path) and
torch bf16 (cuBLAS + elementwise), CUDA-graph replayed."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from massive_marl_benchmark_b200.mlp import FusedMLP
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
dims = [388, 1024, 1024, 512, 80]
mods = []
for i in range(len(dims) - 1):
    mods.append(torch.nn.Linear(dims[i], dims[i + 1]))
    if i < len(dims) - 2: mods.append(torch.nn.ELU())
net = torch.nn.Sequential(*mods).to(dev)
f = FusedMLP.from_sequential(net, dev)
nb = torch.nn.Sequential(*[m for m in net]).to(torch.bfloat16) if False else None
import copy
nb = copy.deepcopy(net).to(torch.bfloat16)
def graph_time(fn, iters=50):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
out = []
for M in (1024, 4096, 16384, 65536):
    x = torch.randn(M, dims[0], device=dev); xb = x.to(torch.bfloat16)
    y = torch.empty(M, dims[-1], device=dev)
    flops = 2 * M * sum(dims[i] * dims[i + 1] for i in range(len(dims) - 1))
    with torch.no_grad():
        t_ours = graph_time(lambda: f(x, y)); t32 = graph_time(lambda: net(x)); t16 = graph_time(lambda: nb(xb))
    row = {"M": M, "ours_ms": t_ours, "ours_tflops": flops / t_ours / 1e9, "torch_fp32_ms": t32, "torch_bf16_ms": t16, "torch_bf16_tflops": flops / t16 / 1e9}
    out.append(row); print({k: round(v, 4) if isinstance(v, float) else v for k, v in row.items()}, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/mlp_batch_sweep.json", "w"), indent=1)
