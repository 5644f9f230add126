#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 8) > gpurun_out/l_pytest.log
tail -n 4 gpurun_out/l_pytest.log
(timeout 300 python tools/bits_bench.py 2>&1 | tail -n 1) > gpurun_out/l_bits.log; cat gpurun_out/l_bits.log
(TC_SWEEP_CASES=0,1,2,3 timeout 900 python tools/env_pack_sweep.py auto 2>&1 | cut -c1-200) > gpurun_out/l_sweep.log; cat gpurun_out/l_sweep.log
(TC_TRACK_GROUP=32 TC_SWEEP_CASES=0 timeout 900 python tools/env_pack_sweep.py auto 2>&1 | cut -c1-200) >> gpurun_out/l_sweep.log; tail -n 1 gpurun_out/l_sweep.log
