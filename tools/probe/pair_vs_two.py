"""Actor + critic forward (M = 4096): one chain launch with grid z = 2 against two chain launches back to back."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from massive_marl_benchmark_b200.mlp import FusedMLP, GroupedMLP  # noqa: E402

dev = torch.device("cuda:0")


def net(dims):
    mods = []
    for i in range(len(dims) - 1):
        mods.append(torch.nn.Linear(dims[i], dims[i + 1]))
        if i < len(dims) - 2:
            mods.append(torch.nn.ELU())
    return torch.nn.Sequential(*mods).to(dev)


def timeit(fn, iters=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for M in (2048, 4096, 8192, 16384):
    a, c = net([388, 1024, 1024, 512, 80]), net([388, 1024, 1024, 512, 80])
    fa, fc = FusedMLP.from_sequential(a, dev), FusedMLP.from_sequential(c, dev)
    pair = GroupedMLP([fa, fc])
    x = torch.randn(M, 388, device=dev)
    print(M, "pair z=2: %.1f us" % timeit(lambda: pair([x, x])), " two launches: %.1f us" % timeit(lambda: (fa(x), fc(x))),
          " one: %.1f us" % timeit(lambda: fa(x)))
