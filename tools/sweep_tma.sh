for s in 0 150 300 450 600; do
  TC_TMA_PERMILLE=$s python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tma_permille',$s, round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['step_share_ms']['raster'],4), round(d['roofline']['frac'],4))"
done
