set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r01f_pytest_gpu.log 2>&1; tail -2 gpurun_out/r01f_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01f_smoke.log 2>&1; tail -2 gpurun_out/r01f_smoke.log
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/r01f_bench_ref.log 2>&1; tail -1 gpurun_out/r01f_bench_ref.log | cut -c1-300
python bench.py > gpurun_out/r01f_bench.log 2>&1; tail -1 gpurun_out/r01f_bench.log | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01f_launches.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-fill-context > gpurun_out/r01f_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_render_classes -s 4 -c 1 -o gpurun_out/r01f_prof_render python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-fill-context > gpurun_out/r01f_ncu_full.log 2>&1
python tools/bench_configs.py > gpurun_out/r01f_configs.log 2>&1; cat gpurun_out/r01f_configs.log
python tools/graph_bench.py > gpurun_out/r01f_graph.log 2>&1; cat gpurun_out/r01f_graph.log
F="warning\|extern\|\^\|Remark\|^$\|Runtime\|ret = \|print("
python tools/timeline.py 16384 simple_layout 84 84 2>&1 | grep -v "$F" > gpurun_out/r01f_tl_simple84.log; python tools/timeline.py 16384 knuffingen 128 160 2>&1 | grep -v "$F" > gpurun_out/r01f_tl_knuff128.log
ncu --set full --clock-control none --import-source on -k regex:tc_render_env_kernel -s 4 -c 1 -o gpurun_out/r01f_prof_env84 python tools/small_prof.py simple_layout 84 84 > gpurun_out/r01f_ncu_env84.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_render_env_kernel -s 4 -c 1 -o gpurun_out/r01f_prof_env128 python tools/small_prof.py knuffingen 128 160 32768 > gpurun_out/r01f_ncu_env128.log 2>&1
