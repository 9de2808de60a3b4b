import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from massive_marl_benchmark_b200.storage import RolloutStorage
dev = torch.device("cuda:0")
T, N = 16, 4194304
st = RolloutStorage(N, T, (1,), (0,), (1,), dev)
st.rewards.normal_(); st.values.normal_()
lv = torch.randn(N, 1, device=dev)
for _ in range(3):
    st.compute_returns_scan(lv, 0.96, 0.95)
torch.cuda.synchronize()
