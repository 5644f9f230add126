#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python tools/env_pack_sweep.py 0 1 2 4 cls 2>&1) > gpurun_out/b_sweep.log
for p in 2 4; do
(TC_ENV_PACK=$p timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch_matches_oracle or odd_resolutions or policy_formats or per_env_params or autoreset or visible_set or grouped" 2>&1 | tail -15) > gpurun_out/b_pytest_pack$p.log
done
(timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/b_pytest_configs.log
cat gpurun_out/b_sweep.log; tail -3 gpurun_out/b_pytest_pack*.log gpurun_out/b_pytest_configs.log
