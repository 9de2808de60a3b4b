"""Per-layer timing of the tcgen05 MLP forward (PPO actor, M = 4096) against cuBLAS bf16 per layer."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from massive_marl_benchmark_b200 import _lib as L
from massive_marl_benchmark_b200.mlp import FusedMLP

dev = torch.device("cuda:0")
dims = [388, 1024, 1024, 512, 80]
M = int(os.environ.get("M", 4096))
mods = []
for i in range(len(dims) - 1):
    mods.append(torch.nn.Linear(dims[i], dims[i + 1]))
    if i < len(dims) - 2: mods.append(torch.nn.ELU())
net = torch.nn.Sequential(*mods).to(dev)
f = FusedMLP.from_sequential(net, dev)
x = torch.randn(M, dims[0], device=dev)
f(x); torch.cuda.synchronize()
Mpad, acts = f._buffers(M)
lib = L.lib()

def timeit(fn, iters=200, warm=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

out = torch.empty(M, dims[-1], device=dev)
for i, l in enumerate(f.layers):
    p = L.MlpLayerParams()
    n_tile = f._n_tile(l, Mpad) if hasattr(f, "_n_tile") else l.n_tile
    p.M, p.N, p.K, p.Mpad, p.Kpad, p.Npad, p.n_tile, p.epilogue = M, l.N, l.K, Mpad, l.Kpad, l.Npad, n_tile, l.epilogue
    p.x, p.w, p.bias = L.ptr(acts[i]), L.ptr(l.w), L.ptr(l.bias)
    if i == len(f.layers) - 1: p.y, p.y_stride = L.ptr(out), out.stride(0)
    else: p.y, p.y_stride = L.ptr(acts[i + 1]), acts[i + 1].stride(0)
    st = L.stream_ptr()
    us = timeit(lambda: lib.mmb_mlp_layer(p, st))
    a = acts[i][:M]; w = l.w
    us_t = timeit(lambda: torch.matmul(a, w.t()))
    fl = 2 * M * l.Kpad * l.Npad
    print("layer %d  K=%4d N=%4d n_tile=%3d ctas=%3d  ours %.1f us (%.0f TF)  cublas-bf16 gemm only %.1f us" % (
        i, l.Kpad, l.Npad, n_tile, (Mpad // 128) * (l.Npad // n_tile), us, fl / us / 1e6, us_t))
us = timeit(lambda: f(x)); print("whole forward, eager %.1f us" % us)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    f(x, out)
us = timeit(lambda: g.replay()); print("whole forward, graph replay %.1f us" % us)
nb = net.to(torch.bfloat16); xb = x.to(torch.bfloat16)
with torch.no_grad():
    nb(xb); torch.cuda.synchronize()
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        yb = nb(xb)
us = timeit(lambda: g2.replay()); print("torch bf16 (cuBLAS + elementwise), graph replay %.1f us" % us)
