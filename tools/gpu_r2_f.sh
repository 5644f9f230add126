#!/bin/bash
mkdir -p gpurun_out
(TC_SWEEP_CASES=2,0,1,4 timeout 900 python tools/env_pack_sweep.py auto 2>&1) > gpurun_out/f_sweep.log
(TC_ENV_CHUNKS=2 TC_SWEEP_CASES=2 timeout 900 python tools/env_pack_sweep.py 2 2>&1) >> gpurun_out/f_sweep.log
cat gpurun_out/f_sweep.log
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 8) > gpurun_out/f_pytest.log
tail -n 4 gpurun_out/f_pytest.log
