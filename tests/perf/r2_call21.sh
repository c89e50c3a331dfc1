#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 200 python tests/perf/encoder_microbench.py 16 > gpurun_out/r2c21_encoder_microbench.json 2> gpurun_out/r2c21_encoder_microbench.err; cat gpurun_out/r2c21_encoder_microbench.json; tail -3 gpurun_out/r2c21_encoder_microbench.err
timeout 200 python tests/perf/encoder_microbench.py 2 > gpurun_out/r2c21_encoder_microbench_T2.json 2>/dev/null; cat gpurun_out/r2c21_encoder_microbench_T2.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c21_bench_n1.json 2> gpurun_out/r2c21_bench_n1.err
tail -3 gpurun_out/r2c21_bench_n1.err
python - <<'P'
import json
l=json.load(open('gpurun_out/r2c21_bench_n1.json')); r=l['roofline']
print('n1', l['value'], l['ms_per_step'], 'lat', l['latency_ms_per_clip'], 'e2e', l['e2e']['value'], 'parity', l['parity_check']['bit_identical'], 'msda us', r['us_per_launch'], 'frac', r['frac'])
print(r['our_kernels_ms_per_clip'])
print(r['our_kernel_launches_per_clip'])
P
