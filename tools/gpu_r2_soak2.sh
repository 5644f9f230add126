#!/bin/bash
# tracking-kernel soak: small frames, many envs and steps, each tracking variant forced in turn (thread per env, 8 and 32 lanes per env)
mkdir -p gpurun_out
(TC_TRACK_MODE=thread timeout 900 python tools/parity_soak.py 8192 150 32 48 knuffingen | tail -n 1) > gpurun_out/r02_soak_track_thread.json
(TC_TRACK_MODE=warp TC_TRACK_GROUP=8 timeout 900 python tools/parity_soak.py 8192 150 32 48 knuffingen | tail -n 1) > gpurun_out/r02_soak_track_g8.json
(TC_TRACK_MODE=warp TC_TRACK_GROUP=32 timeout 900 python tools/parity_soak.py 4096 100 32 48 simple_layout | tail -n 1) > gpurun_out/r02_soak_track_g32.json
cut -c1-330 gpurun_out/r02_soak_track_*.json
