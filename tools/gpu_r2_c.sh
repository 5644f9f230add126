#!/bin/bash
mkdir -p gpurun_out
echo "--- track warp" > gpurun_out/c_sweep.log
(TC_TRACK_MODE=warp TC_SWEEP_CASES=0,2 timeout 600 python tools/env_pack_sweep.py 0 2>&1) >> gpurun_out/c_sweep.log
echo "--- track thread" >> gpurun_out/c_sweep.log
(TC_TRACK_MODE=thread TC_SWEEP_CASES=0,2 timeout 600 python tools/env_pack_sweep.py 0 2>&1) >> gpurun_out/c_sweep.log
echo "--- equal occupancy experiment" >> gpurun_out/c_sweep.log
(TC_SWEEP_CASES=4 timeout 600 python tools/env_pack_sweep.py 0 1 2 4 2>&1) >> gpurun_out/c_sweep.log
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 15) > gpurun_out/c_pytest.log
cat gpurun_out/c_sweep.log; tail -n 5 gpurun_out/c_pytest.log
