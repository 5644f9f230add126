#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/h_bits_launches.csv python tools/bits_bench.py > gpurun_out/h_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/h_bits_launches.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
kn, mv = H.index("Kernel Name"), H.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > mv:
        agg[r[kn][:60]].append(float(r[mv].replace(",", "")))
for k, v in agg.items():
    print(f"{k:60s} n={len(v):3d} mean {sum(v) / len(v) / 1e3:9.1f} us  min {min(v) / 1e3:9.1f}")
PY
