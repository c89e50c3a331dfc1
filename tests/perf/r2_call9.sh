#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/r2c9_gpu_tests.log 2>&1
tail -6 gpurun_out/r2c9_gpu_tests.log
timeout 300 python tests/perf/msda_microbench.py > gpurun_out/r2c9_msda_microbench.log 2>&1
grep "frames': 16" gpurun_out/r2c9_msda_microbench.log | grep -v reference | cut -c1-160
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c9_bench_n1.json 2> gpurun_out/r2c9_bench_n1.err
python - <<'P'
import json
l=json.load(open('gpurun_out/r2c9_bench_n1.json')); r=l['roofline']
print(l['value'], l['ms_per_step'], 'lat', l['latency_ms_per_clip'], 'e2e', l['e2e']['value'], 'parity', l['parity_check']['bit_identical'], 'msda us', r['us_per_launch'], 'frac', r['frac'])
print(r['our_kernels_ms_per_clip'])
P
for c in 2 3; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config $c > gpurun_out/r2c9_bench_config$c.json 2> gpurun_out/r2c9_bench_config$c.err
  cut -c1-330 gpurun_out/r2c9_bench_config$c.json; tail -2 gpurun_out/r2c9_bench_config$c.err
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:msda_fwd_staged -c 1 -o gpurun_out/r2c9_ncu_msda -f \
  python tests/perf/msda_profile_target.py 8 > gpurun_out/r2c9_ncu_msda.log 2>&1
tail -2 gpurun_out/r2c9_ncu_msda.log
