#!/bin/bash
mkdir -p gpurun_out
for c in ours library; do
  if [ $c = ours ]; then unset DVIS_BENCH_TRACKER_LIBRARY_ATTENTION; else export DVIS_BENCH_TRACKER_LIBRARY_ATTENTION=1; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c34_bench_tracker_attn_$c.json 2> gpurun_out/r2c34_$c.err
  python - <<P
import json
l=json.load(open('gpurun_out/r2c34_bench_tracker_attn_$c.json')); r=l['roofline']
print('tracker attention $c:', l['value'], l['ms_per_step'], 'lat', l['latency_ms_per_clip'], 'e2e', l['e2e']['value'])
P
done
timeout 300 python -m pytest tests/test_linear_tc_gpu.py -q -x -k partition -p no:cacheprovider 2>&1 | tail -2
