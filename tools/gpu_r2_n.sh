#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python tools/bits_bench.py 2>&1 | tail -n 1) > gpurun_out/n.log
(TC_SWEEP_CASES=0,1,2 timeout 900 python tools/env_pack_sweep.py auto 2>&1 | cut -c1-170) >> gpurun_out/n.log; cat gpurun_out/n.log
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -k "batch_matches or odd_resolutions or guard_bands or grouped or policy_formats or banded or config4" 2>&1 | tail -n 3) > gpurun_out/n_pytest.log; tail -n 2 gpurun_out/n_pytest.log
