#!/bin/bash
# round-2 profiling pass: per BASELINE config the ncu launch list of the bench command and one --set full capture of its render kernel(s)
mkdir -p gpurun_out
B="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-fill-context --min-seconds 0"
for c in 3 2 1 4 5; do
  python bench.py --config $c $B > gpurun_out/p_plain_c$c.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_c$c.csv python bench.py --config $c $B > gpurun_out/p_ncu_l_c$c.log 2>&1
  python tools/ncu_summary.py launches gpurun_out/r02_launches_c$c.csv > gpurun_out/r02_launches_c$c.txt 2>&1
done
full() {  # name, kernel regex, skip, count, command...
  name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  "$@" > gpurun_out/p_plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o gpurun_out/r02_prof_$name "$@" > gpurun_out/p_ncu_f_$name.log 2>&1
  python tools/ncu_summary.py full gpurun_out/r02_prof_$name.ncu-rep > gpurun_out/r02_ncu_$name.txt 2>&1
}
full c3 tc_render_classes 4 1 python bench.py --config 3 $B
full c2 tc_render_envs 4 1 python bench.py --config 2 $B
full c1 tc_render_env_banded 4 1 python bench.py --config 1 $B
full c5 tc_render 12 3 python bench.py --config 5 $B
full bits "tc_prims|tc_draw_class" 8 2 python tools/bits_bench.py
full knuff128 tc_render_envs 6 1 python tools/small_prof.py knuffingen 128 160 32768
full track tc_track_thread 6 1 python tools/small_prof.py knuffingen 128 160 32768
ls -la gpurun_out/*.ncu-rep
rm -f gpurun_out/r02_prof_c1.ncu-rep gpurun_out/r02_prof_c5.ncu-rep gpurun_out/r02_prof_track.ncu-rep   # keep the transfer small: their summaries stay
head -30 gpurun_out/r02_launches_c3.txt
