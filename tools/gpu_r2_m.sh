#!/bin/bash
mkdir -p gpurun_out
(TC_SWEEP_CASES=0,1,2,3 timeout 900 python tools/env_pack_sweep.py auto 2>&1 | cut -c1-200) > gpurun_out/m_sweep.log; cat gpurun_out/m_sweep.log
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -k "batch_matches or odd_resolutions or guard_bands or grouped or policy_formats" 2>&1 | tail -n 5) > gpurun_out/m_pytest.log; tail -n 3 gpurun_out/m_pytest.log
