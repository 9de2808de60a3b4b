import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from massive_marl_benchmark_b200 import _lib as L
from massive_marl_benchmark_b200.mlp import FusedMLP
dev = torch.device("cuda:0")
net = torch.nn.Sequential(torch.nn.Linear(1024, 1024), torch.nn.ELU(), torch.nn.Linear(1024, 1024)).to(dev)
f = FusedMLP.from_sequential(net, dev)
x = torch.randn(4096, 1024, device=dev)
y = f(x); torch.cuda.synchronize()
st = (C.c_uint32 * 4)(); L.lib().mmb_mlp_debug_status(st); print("debug status", [hex(v) for v in st])
ref = net(x).detach()
print("max abs err", float((y - ref).abs().max()), "ref max", float(ref.abs().max()))
