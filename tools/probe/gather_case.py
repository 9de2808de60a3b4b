"""One case for ncu: shuffle + gather of 16 M transitions (OneAnt-width payload, 356 B per transition), fused epoch gather with
shuffle_group 8 and 1 (tools/sweep_storage.py's middle size).  A few launches only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from massive_marl_benchmark_b200.storage import RolloutStorage  # noqa: E402

dev = torch.device("cuda:0")
T, N = 16, 1048576
st = RolloutStorage(N, T, (60,), (0,), (8,), dev, "random")
st.observations.normal_(); st.actions.normal_(); st.mu.normal_(); st.sigma.normal_(); st.values.normal_(); st.returns.normal_()
bufs = None
for group in (8, 1):
    st.shuffle_group = group
    for _ in range(2):
        st.new_epoch()
        for k in range(4):
            bufs = st.gather_epoch_minibatch(k, 4, bufs)
torch.cuda.synchronize()
print("ok")
