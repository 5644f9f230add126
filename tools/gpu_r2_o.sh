#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4) > gpurun_out/o_pytest.log; tail -n 2 gpurun_out/o_pytest.log
timeout 600 python bench.py --config 2 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/o_c2.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/o_c2.log").read().strip().splitlines()[-1])
print("config 2: value", round(d["value"] / 1e6, 2), "eager", round(d["eager"]["value"] / 1e6, 2), "sustained", round(d["sustained"]["value"] / 1e6, 2), "e2e", round(d["e2e"]["value"] / 1e6, 2))
PY
