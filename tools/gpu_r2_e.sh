#!/bin/bash
mkdir -p gpurun_out
(TC_SWEEP_CASES=2,0,1,3,4 timeout 900 python tools/env_pack_sweep.py auto 0 2>&1) > gpurun_out/e_sweep.log
(TC_ENV_CHUNKS=1 TC_SWEEP_CASES=0,1 timeout 900 python tools/env_pack_sweep.py 2 2>&1) >> gpurun_out/e_sweep.log
cat gpurun_out/e_sweep.log
