#!/bin/bash
# the round's soaks once more on the final library (shorter step counts: ~6 minutes)
mkdir -p gpurun_out
(timeout 900 python tools/parity_soak.py 4096 100 128 160 knuffingen | tail -n 1) > gpurun_out/r02_soak_knuff128.json
(timeout 900 python tools/parity_soak.py 4096 100 84 84 simple_layout | tail -n 1) > gpurun_out/r02_soak_simple84.json
(timeout 900 python tools/parity_soak.py 1024 60 240 320 knuffingen | tail -n 1) > gpurun_out/r02_soak_knuff240.json
(TC_FMT=classes_bits timeout 900 python tools/parity_soak.py 512 80 480 640 knuffingen | tail -n 1) > gpurun_out/r02_soak_knuff480_bits.json
(timeout 900 python tools/parity_soak.py 384 80 480 640 knuffingen | tail -n 1) > gpurun_out/r02_soak_knuff480.json
(TC_TRACK_MODE=thread timeout 900 python tools/parity_soak.py 8192 150 32 48 knuffingen | tail -n 1) > gpurun_out/r02_soak_track_thread.json
(TC_TRACK_MODE=warp TC_TRACK_GROUP=8 timeout 900 python tools/parity_soak.py 8192 150 32 48 knuffingen | tail -n 1) > gpurun_out/r02_soak_track_g8.json
(TC_TRACK_MODE=warp TC_TRACK_GROUP=32 timeout 900 python tools/parity_soak.py 4096 100 32 48 simple_layout | tail -n 1) > gpurun_out/r02_soak_track_g32.json
python - <<'PY'
import json, glob
tot = 0
for f in sorted(glob.glob("gpurun_out/r02_soak_*.json")):
    d = json.load(open(f))
    tot += d["env_steps"]
    print(f.split("/")[-1], d["env_steps"], "bad frames/px/idx/flags", d["bad_frames"], d["bad_pixels"], d["bad_index_rows"], d["bad_flags"], "resets", d["resets"], d.get("library_so_hash"))
print("total env-steps", tot)
PY
