#!/bin/bash
mkdir -p gpurun_out
for gs in 1 5; do
timeout 300 python bench.py --config 2 --no-cpu-baseline --no-e2e --min-seconds 0 --graph-steps $gs 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('graph-steps $gs: value', round(d['value']/1e6,2), 'eager', round(d['eager']['value']/1e6,2), d['cuda_graph']['note'][-40:])"
done
