#!/bin/bash
mkdir -p gpurun_out
python tools/bits_bench.py > gpurun_out/q_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_bits.csv python tools/bits_bench.py > gpurun_out/q_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/r02_launches_bits.csv | head -12
tail -n 1 gpurun_out/q_plain.log
