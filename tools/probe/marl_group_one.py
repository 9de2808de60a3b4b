"""A few grouped MARL team forwards (ten actors, M = 4096) - the command the ncu launch list / `--set full` is pointed at."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from massive_marl_benchmark_b200.mlp import FusedMLP, GroupedMLP
dev = torch.device("cuda:0")
M, G = 4096, 10
gen = torch.Generator().manual_seed(0)


def sd(in_dim, out_dim, head):
    d = {"base.feature_norm.weight": torch.ones(in_dim), "base.feature_norm.bias": torch.zeros(in_dim)}
    dims = [("base.mlp.fc1", in_dim), ("base.mlp.fc2.0", 512), ("base.mlp.fc2.1", 512)]
    for n, k in dims:
        d[n + ".0.weight"] = torch.randn(512, k, generator=gen) * (1.4 / k ** 0.5)
        d[n + ".0.bias"] = torch.zeros(512)
        d[n + ".2.weight"] = torch.ones(512)
        d[n + ".2.bias"] = torch.zeros(512)
    d[head + ".weight"] = torch.randn(out_dim, 512, generator=gen) * 0.01
    d[head + ".bias"] = torch.zeros(out_dim)
    return d


kind = sys.argv[1] if len(sys.argv) > 1 else "actors"
in_dim, out_dim, head = (46, 8, "act.action_out.fc_mean") if kind == "actors" else (388, 1, "v_out")
team = GroupedMLP([FusedMLP.from_marl_state_dict(sd(in_dim, out_dim, head), head, dev) for _ in range(G)])
xs = torch.randn(G, M, in_dim, device=dev)
for _ in range(4):
    out = team(xs)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
