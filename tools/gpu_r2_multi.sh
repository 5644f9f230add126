#!/bin/bash
# multi-GPU pass on one 8-GPU box: config 5 (strong scaling, resolution groups) on 2/4/8, configs 4 and 3 on 8
mkdir -p gpurun_out
run() { n=$1; c=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n + 10 * c)) bench.py --gpus $n --config $c --steps 50 --warmup 5 --min-seconds 2 --no-cpu-baseline "$@" 2> gpurun_out/m_c${c}_n${n}.err | grep '^{' > gpurun_out/r02_bench_c${c}_${n}gpu.log
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_c${c}_${n}gpu.log").read().strip().splitlines()[-1])
    print("config $c x $n GPUs:", round(d["value"] / 1e6, 2), "M env-steps/s, sustained", round(d["sustained"]["value"] / 1e6, 2), "e2e", round(d["e2e"]["value"] / 1e6, 2), d["config"]["envs_per_gpu"], "envs/GPU")
except Exception as e:
    print("config $c x $n GPUs: FAILED", e)
PY
}
nvidia-smi -L | wc -l
run 2 5; run 4 5; run 8 5; run 8 4; run 8 2
