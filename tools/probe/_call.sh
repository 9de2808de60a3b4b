timeout 300 python -m pytest tests/test_gpu_ppo_rollout.py tests/test_gpu_graphed_step.py -x -q -m gpu > gpurun_out/r2s4_pytest_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s4_pytest_d.log
tail -15 gpurun_out/r2s4_pytest_d.log
timeout 200 python tools/bench_ppo_rollout.py > gpurun_out/r2s4_ppo_rollout_d.log 2>&1; tail -6 gpurun_out/r2s4_ppo_rollout_d.log
