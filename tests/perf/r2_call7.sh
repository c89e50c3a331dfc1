#!/bin/bash
# round 2, GPU call 7: full suite (incl. config-5 clip tests), the new bench.py (default, fp32, configs 2 / 3), MSDA ncu with source
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider --durations=8 > gpurun_out/r2c7_gpu_tests.log 2>&1
tail -25 gpurun_out/r2c7_gpu_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c7_bench_n1.json 2> gpurun_out/r2c7_bench_n1.err
cat gpurun_out/r2c7_bench_n1.json; tail -3 gpurun_out/r2c7_bench_n1.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision fp32 > gpurun_out/r2c7_bench_n1_fp32.json 2> gpurun_out/r2c7_bench_n1_fp32.err
cat gpurun_out/r2c7_bench_n1_fp32.json; tail -3 gpurun_out/r2c7_bench_n1_fp32.err
for c in 2 3; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config $c > gpurun_out/r2c7_bench_config$c.json 2> gpurun_out/r2c7_bench_config$c.err
  cat gpurun_out/r2c7_bench_config$c.json; tail -3 gpurun_out/r2c7_bench_config$c.err
done
timeout 200 python tests/perf/pipeline_breakdown.py 16 > gpurun_out/r2c7_breakdown.log 2>&1; tail -1 gpurun_out/r2c7_breakdown.log
timeout 200 python tests/perf/kernel_table.py 16 > /dev/null 2>&1; cp gpurun_out/kernel_table.txt gpurun_out/r2c7_kernel_table_T16.txt; head -30 gpurun_out/r2c7_kernel_table_T16.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:msda_fwd_staged -c 2 -o gpurun_out/r2c7_ncu_msda -f \
  python tests/perf/msda_profile_target.py 8 > gpurun_out/r2c7_ncu_msda.log 2>&1
tail -3 gpurun_out/r2c7_ncu_msda.log
