#!/bin/bash
# round-2 final single-GPU evidence: GPU tests, smoke, every BASELINE config as a full bench line, the CPU arms
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 6) > gpurun_out/r02_pytest_gpu.log; tail -n 2 gpurun_out/r02_pytest_gpu.log
(timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2) > gpurun_out/r02_smoke.log; cat gpurun_out/r02_smoke.log
for c in 3 2 1 4 5; do
  timeout 900 python bench.py --config $c 2> gpurun_out/f_bench_c$c.err | grep '^{' > gpurun_out/r02_bench_c${c}_1gpu.log
done
timeout 600 python bench.py --impl reference --config 3 --steps 100 --warmup 5 2> gpurun_out/f_ref_c3.err | grep '^{' > gpurun_out/r02_bench_reference_arm_c3.log
timeout 600 python bench.py --impl reference --config 1 --steps 100 --warmup 5 2> gpurun_out/f_ref_c1.err | grep '^{' > gpurun_out/r02_bench_reference_arm_c1.log
python - <<'PY'
import json
for c in (3, 2, 1, 4, 5):
    try:
        d = json.loads(open(f"gpurun_out/r02_bench_c{c}_1gpu.log").read().strip().splitlines()[-1])
        cb = d["cpu_baseline"] or {}
        print(c, round(d["value"] / 1e6, 2), "M | sustained", round(d["sustained"]["value"] / 1e6, 2), "| e2e", round(d["e2e"]["value"] / 1e6, 2), "| frac", round(d["roofline"]["frac"], 3),
              "| graph", d["cuda_graph"] and round(d["cuda_graph"]["value"] / 1e6, 2), "| obs->host", d["e2e_obs_to_host"] and d["e2e_obs_to_host"].get("value"),
              "| ref", cb.get("kind"), round(cb.get("value", 0)), "single", cb.get("single_process", {}).get("value"), "| port", d["cpu_baseline_port"] and round(d["cpu_baseline_port"]["value"]))
    except Exception as e:
        print(c, "FAILED", e)
for n in ("c3", "c1"):
    try:
        d = json.loads(open(f"gpurun_out/r02_bench_reference_arm_{n}.log").read().strip().splitlines()[-1])
        print("reference arm", n, round(d["value"], 1), d["cpu_baseline"]["cores"], "cores; single", d["cpu_baseline"]["single_process"]["value"], "; port", d["cpu_baseline_port"]["value"])
    except Exception as e:
        print("reference arm", n, "FAILED", e)
PY
