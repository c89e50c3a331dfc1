#!/bin/bash
# round 2, GPU call 8: MSDA rewrite (FHFMA gather loop, 1-thread softmax, dense phase 1), masked attention prefetch, graphed configs 2/3
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/r2c8_gpu_tests.log 2>&1
tail -6 gpurun_out/r2c8_gpu_tests.log
timeout 300 python tests/perf/msda_microbench.py > gpurun_out/r2c8_msda_microbench.log 2>&1
tail -40 gpurun_out/r2c8_msda_microbench.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c8_bench_n1.json 2> gpurun_out/r2c8_bench_n1.err
cat gpurun_out/r2c8_bench_n1.json; tail -3 gpurun_out/r2c8_bench_n1.err
for c in 2 3; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config $c > gpurun_out/r2c8_bench_config$c.json 2> gpurun_out/r2c8_bench_config$c.err
  cat gpurun_out/r2c8_bench_config$c.json; tail -3 gpurun_out/r2c8_bench_config$c.err
done
timeout 300 python tests/perf/temporal_microbench.py > gpurun_out/r2c8_temporal_microbench.log 2>&1
grep -A3 "masked_cross" gpurun_out/r2c8_temporal_microbench.log | head -20
timeout 200 python tests/perf/pipeline_breakdown.py 16 > gpurun_out/r2c8_breakdown.log 2>&1; tail -1 gpurun_out/r2c8_breakdown.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:msda_fwd_staged -c 1 -o gpurun_out/r2c8_ncu_msda -f \
  python tests/perf/msda_profile_target.py 8 > gpurun_out/r2c8_ncu_msda.log 2>&1
tail -2 gpurun_out/r2c8_ncu_msda.log
