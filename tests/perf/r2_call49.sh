#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_temporal_kernels_gpu.py tests/test_configs_gpu.py tests/test_modules_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -2
timeout 400 python tests/perf/per_rank_stage_probe.py 2>&1 | tail -5
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c49_bench_n1.json 2> gpurun_out/r2c49_n1.err
python - <<'P'
import json
l=json.load(open('gpurun_out/r2c49_bench_n1.json')); r=l['roofline']
print('n1', l['value'], l['ms_per_step'], 'lat', l['latency_ms_per_clip'], 'e2e', l['e2e']['value'], 'parity', l['parity_check']['bit_identical'], 'msda us', r['us_per_launch'])
print(r['our_kernels_ms_per_clip'])
P
