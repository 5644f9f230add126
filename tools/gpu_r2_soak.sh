#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python tools/parity_soak.py 4096 150 128 160 knuffingen | tail -n 1) > gpurun_out/r02_soak_knuff128.json
(timeout 900 python tools/parity_soak.py 4096 150 84 84 simple_layout | tail -n 1) > gpurun_out/r02_soak_simple84.json
(timeout 900 python tools/parity_soak.py 1024 100 240 320 knuffingen | tail -n 1) > gpurun_out/r02_soak_knuff240.json
(TC_FMT=classes_bits timeout 900 python tools/parity_soak.py 512 120 480 640 knuffingen | tail -n 1) > gpurun_out/r02_soak_knuff480_bits.json
(timeout 900 python tools/parity_soak.py 384 120 480 640 knuffingen | tail -n 1) > gpurun_out/r02_soak_knuff480.json
cut -c1-400 gpurun_out/r02_soak_*.json
