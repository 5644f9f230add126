#!/bin/bash
mkdir -p gpurun_out
(TC_PRIMS_PATH=1 timeout 300 python tools/bits_bench.py 2>&1 | tail -n 2) > gpurun_out/g_bits.log
(TC_PRIMS_PATH=0 timeout 300 python tools/bits_bench.py 2>&1 | tail -n 2) >> gpurun_out/g_bits.log
cat gpurun_out/g_bits.log
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "policy_formats or banded or full_size" 2>&1 | tail -n 8) > gpurun_out/g_pytest.log
tail -n 4 gpurun_out/g_pytest.log
