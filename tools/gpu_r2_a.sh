#!/bin/bash
# round-2 first GPU pass: full GPU test suite, every BASELINE config once, the reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/a_pytest.log
for c in 3 2 1 4 5; do
  (timeout 400 python bench.py --config $c --steps 50 --warmup 5 --cpu-seconds 6 2>&1 | tail -5) > gpurun_out/a_bench_c$c.log
done
(timeout 400 python bench.py --impl reference --config 3 --steps 20 --warmup 2 2>&1 | tail -3) > gpurun_out/a_ref_c3.log
tail -3 gpurun_out/a_pytest.log
