#!/bin/bash
mkdir -p gpurun_out
for c in none 100 70 85; do
  if [ $c = none ]; then unset DVIS_SMEM_CARVEOUT; else export DVIS_SMEM_CARVEOUT=$c; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c33_bench_carveout_$c.json 2> gpurun_out/r2c33_$c.err
  python - <<P
import json
l=json.load(open('gpurun_out/r2c33_bench_carveout_$c.json')); r=l['roofline']
print('carveout $c:', l['value'], l['ms_per_step'], 'lat', l['latency_ms_per_clip'], 'e2e', l['e2e']['value'], 'msda us', r['us_per_launch'])
P
done
