"""Timeline inside one `mmb_mlp_chain` launch (PPO actor, M = 4096): %globaltimer stamps of the first eight CTAs
(two clusters), printed per layer relative to the kernel's first stamp.  Run with MMB_CHAIN_TRACE=1."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

os.environ.setdefault("MMB_CHAIN_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from massive_marl_benchmark_b200 import _lib as L  # noqa: E402
from massive_marl_benchmark_b200.mlp import FusedMLP  # noqa: E402

dev = torch.device("cuda:0")
dims = [388, 1024, 1024, 512, 80]
mods = []
for i in range(len(dims) - 1):
    mods.append(torch.nn.Linear(dims[i], dims[i + 1]))
    if i < len(dims) - 2:
        mods.append(torch.nn.ELU())
net = torch.nn.Sequential(*mods).to(dev)
mlp = FusedMLP.from_sequential(net, dev)
x = torch.randn(4096, dims[0], device=dev)
for _ in range(20):
    mlp(x)
torch.cuda.synchronize()
words = 8 * 6 * 24
buf = (C.c_uint64 * words)()
L.check(L.lib().mmb_mlp_chain_trace(buf, words), "mmb_mlp_chain_trace")
allw = np.array(buf[:], dtype=np.int64)
t = allw[:8 * 6 * 8].reshape(8, 6, 8)
kbt = allw[8 * 6 * 8:].reshape(8, 6, 16)
t0 = t[t > 0].min()
names = ["begin", "first_operands", "mma_issued", "acc_complete", "computed", "landed", "boundary"]
out = {}
for cta in range(8):
    rows = []
    for l in range(len(dims) - 1):
        rows.append({n: (int(t[cta, l, e] - t0) if t[cta, l, e] > 0 else None) for e, n in enumerate(names)})
    out["cta%d" % cta] = rows
for l in range(len(dims) - 1):
    print("layer %d" % l, {n: [out["cta%d" % c][l][n] for c in (0, 1, 4)] for n in names})
for l in range(len(dims) - 1):
    print("layer %d k-block operand arrival, cta0:" % l, [int(v - t0) for v in kbt[0, l] if v > 0])
    out["cta0_layer%d_kblock_ns" % l] = [int(v - t0) for v in kbt[0, l] if v > 0]
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/chain_trace.json", "w"), indent=1)
