#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6
timeout 300 python tests/perf/mask_gemm_microbench.py > gpurun_out/r2c32_mask_gemm_microbench.log 2>&1; grep "ours" gpurun_out/r2c32_mask_gemm_microbench.log | cut -c1-200
cp gpurun_out/mask_gemm_microbench.json gpurun_out/r2c32_mask_gemm_microbench.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c32_bench_n1.json 2> gpurun_out/r2c32_bench_n1.err
tail -3 gpurun_out/r2c32_bench_n1.err
python - <<'P'
import json
l=json.load(open('gpurun_out/r2c32_bench_n1.json')); r=l['roofline']
print('n1', l['value'], l['ms_per_step'], 'lat', l['latency_ms_per_clip'], 'e2e', l['e2e']['value'], 'parity', l['parity_check']['bit_identical'], 'msda us', r['us_per_launch'], 'frac', r['frac'])
print(r['our_kernels_ms_per_clip'])
P
