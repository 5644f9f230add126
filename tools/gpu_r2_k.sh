#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 12) > gpurun_out/k_pytest.log
tail -n 5 gpurun_out/k_pytest.log
(TC_SWEEP_CASES=3 timeout 300 python tools/env_pack_sweep.py auto cls 2>&1) > gpurun_out/k_sweep.log; cat gpurun_out/k_sweep.log
