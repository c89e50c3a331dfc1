#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_mask_gemm_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -8
timeout 200 python tests/perf/fp32_tier_errors.py 2>&1 | tail -30 | tee gpurun_out/r2c30_fp32_tier_errors.json
timeout 600 python -m pytest tests/test_modules_gpu.py tests/test_configs_gpu.py tests/test_postprocess_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -5
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --precision fp32 > gpurun_out/r2c30_bench_n1_fp32.json 2> gpurun_out/r2c30_bench_n1_fp32.err
tail -3 gpurun_out/r2c30_bench_n1_fp32.err
python - <<'P'
import json
l=json.load(open('gpurun_out/r2c30_bench_n1_fp32.json')); r=l['roofline']
print('fp32 n1', l['value'], l['ms_per_step'], 'e2e', l['e2e']['value'], 'parity', l['parity_check'])
print(r and r['our_kernels_ms_per_clip'])
P
