#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python tools/sanitize_cases.py 2>&1 | tail -n 25) > gpurun_out/j_cases_plain.log
tail -n 3 gpurun_out/j_cases_plain.log
(timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -x -q -k "camera_randomisation or visible_set or per_env_params or grouped or replays_reference_trace and knuff_camrand" 2>&1 | tail -n 8) > gpurun_out/j_pytest.log
tail -n 4 gpurun_out/j_pytest.log
for tool in memcheck racecheck initcheck synccheck; do
  (TC_SAN_STEPS=2 timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py 2>&1 | tail -n 60) > gpurun_out/j_san_$tool.log
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_cases:|Error|hazard" gpurun_out/j_san_$tool.log | head -8
done
