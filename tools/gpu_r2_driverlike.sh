#!/bin/bash
# what the driver does at round end, at 2 GPUs: the reference arm and this repo's arm under torch.distributed.run
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --impl reference --gpus 2 --steps 20 --warmup 2 2> gpurun_out/d_ref.err | grep '^{' > gpurun_out/d_ref_2gpu.log; echo "reference arm rc=$? lines=$(wc -l < gpurun_out/d_ref_2gpu.log)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 30 --warmup 3 2> gpurun_out/d_ours.err | grep '^{' > gpurun_out/d_ours_2gpu.log; echo "ours rc=$? lines=$(wc -l < gpurun_out/d_ours_2gpu.log)"
python - <<'PY'
import json
a = json.loads(open("gpurun_out/d_ref_2gpu.log").read().strip().splitlines()[-1])
b = json.loads(open("gpurun_out/d_ours_2gpu.log").read().strip().splitlines()[-1])
print("reference", round(a["value"], 1), a["impl"], a["n_gpus"], "| ours", round(b["value"] / 1e6, 2), "M", b["n_gpus"], "| same config:", a["config"] == b["config"], "| metric same:", a["metric"] == b["metric"], "| traffic", b["roofline"]["traffic"] is not None, "| cpu_baseline in ours at N=2:", b["cpu_baseline"])
PY
