timeout 300 python -m pytest tests/test_gpu_ppo_rollout.py -x -q -m gpu > gpurun_out/r2s4_pytest_g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s4_pytest_g.log
tail -12 gpurun_out/r2s4_pytest_g.log
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-rollouts 0 > gpurun_out/r2s4_bench20_b.json 2> gpurun_out/r2s4_bench20_b.err; echo "bench rc=$?"; tail -3 gpurun_out/r2s4_bench20_b.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2s4_bench20_b.json")); print(d["value"], d["ms_per_step"], d["mlp_forward"]["ms_graph_replay"], d["ppo_rollout"])
PY
