#!/bin/bash
# round 2, GPU call 1: baseline suite, leg-difference diagnosis, kernel table + ncu launch list of the final pipeline
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/r2c1_gpu_tests.log 2>&1
tail -5 gpurun_out/r2c1_gpu_tests.log
timeout 400 python tests/perf/diagnose_legs.py > gpurun_out/r2c1_diagnose_legs.json 2> gpurun_out/r2c1_diagnose_legs.err
cat gpurun_out/r2c1_diagnose_legs.json; tail -5 gpurun_out/r2c1_diagnose_legs.err
timeout 200 python tests/perf/kernel_table.py 16 > /dev/null 2>&1; cp gpurun_out/kernel_table.txt gpurun_out/r2c1_kernel_table_T16.txt
timeout 200 python tests/perf/pipeline_breakdown.py 16 > gpurun_out/r2c1_breakdown.log 2>&1; tail -3 gpurun_out/r2c1_breakdown.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r2c1_launches_final.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eager > gpurun_out/r2c1_bench_under_ncu.log 2>&1
ls -la gpurun_out
