#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python tools/bits_bench.py 2>&1 | tail -n 1) > gpurun_out/i.log
(timeout 300 python tools/rgb_bench.py 2>&1 | tail -n 1) >> gpurun_out/i.log
(TC_LIB=$PWD/tinycarlo_b200/lib/alt_envb3.so timeout 300 python tools/rgb_bench.py 2>&1 | tail -n 1) >> gpurun_out/i.log
(TC_LIB=$PWD/tinycarlo_b200/lib/alt_envb3.so TC_PRIMS_PATH=0 timeout 300 python tools/bits_bench.py 2>&1 | tail -n 1) >> gpurun_out/i.log
cat gpurun_out/i.log
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "policy_formats or banded or odd_resolutions or batch_matches" 2>&1 | tail -n 8) > gpurun_out/i_pytest.log
tail -n 4 gpurun_out/i_pytest.log
