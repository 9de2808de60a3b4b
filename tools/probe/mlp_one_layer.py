"""One 1024 x 1024 layer at M = 4096, a few launches (for an ncu capture of mlp_layer_ws_kernel)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from massive_marl_benchmark_b200.mlp import FusedMLP
dev = torch.device("cuda:0")
net = torch.nn.Sequential(torch.nn.Linear(1024, 1024), torch.nn.ELU(), torch.nn.Linear(1024, 1024), torch.nn.ELU(), torch.nn.Linear(1024, 1024)).to(dev)
f = FusedMLP.from_sequential(net, dev)
x = torch.randn(4096, 1024, device=dev)
for _ in range(4):
    f(x)
torch.cuda.synchronize()
