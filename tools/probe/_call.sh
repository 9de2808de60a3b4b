timeout 300 python -m pytest tests/test_gpu_ppo_rollout.py tests/test_gpu_storage.py tests/test_gpu_ppo_update.py -x -q -m gpu > gpurun_out/r2s4_pytest_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s4_pytest_c.log
tail -5 gpurun_out/r2s4_pytest_c.log
timeout 200 python tools/bench_ppo_rollout.py > gpurun_out/r2s4_ppo_rollout_c.log 2>&1; tail -6 gpurun_out/r2s4_ppo_rollout_c.log
ROLLOUTS=4 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/rollout_launches.csv python tools/probe/rollout_launches.py > gpurun_out/rollout_ncu.log 2>&1
