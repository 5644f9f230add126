for s in 0 3000 6000 9000 14000; do
  TC_STAGGER_NS=$s python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('stagger',$s, round(d['value']), d['ms_per_step'], d['roofline']['step_share']['raster'], d['roofline']['frac'])"
done
