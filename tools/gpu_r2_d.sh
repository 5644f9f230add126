#!/bin/bash
mkdir -p gpurun_out
echo "--- main lib (64 regs, spills), packs" > gpurun_out/d_sweep.log
(TC_SWEEP_CASES=4,2,0 timeout 600 python tools/env_pack_sweep.py 0 1 2 2>&1) >> gpurun_out/d_sweep.log
echo "--- alt lib: packed kernel at 3 blocks/SM, 80 regs, no spills" >> gpurun_out/d_sweep.log
(TC_LIB=$PWD/tinycarlo_b200/lib/alt_minb3.so TC_SWEEP_CASES=4,2,0 timeout 600 python tools/env_pack_sweep.py 1 2 4 2>&1) >> gpurun_out/d_sweep.log
cat gpurun_out/d_sweep.log
