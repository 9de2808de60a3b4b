"""A few launches of the dual-network chain (PPO actor + critic, M = 4096) - the command `ncu --set full -k regex:mlp_chain_duo`
is pointed at (tools/capture_profiles.sh recipe; one GPU, after the plain run exited 0)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from massive_marl_benchmark_b200.mlp import FusedMLP, GroupedMLP
dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096


def net(dims):
    mods = []
    for i in range(len(dims) - 1):
        mods.append(torch.nn.Linear(dims[i], dims[i + 1]))
        if i < len(dims) - 2:
            mods.append(torch.nn.ELU())
    return torch.nn.Sequential(*mods).to(dev)


dims = [388, 1024, 1024, 512, 80]
pair = GroupedMLP([FusedMLP.from_sequential(net(dims), dev), FusedMLP.from_sequential(net(dims), dev)])
x = torch.randn(M, dims[0], device=dev)
for _ in range(12):
    out = pair([x, x])
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
