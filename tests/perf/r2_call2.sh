#!/bin/bash
# round 2, GPU call 2: new temporal-stage kernels -- parity tests, micro-benchmarks, determinism re-check, bench line
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_temporal_kernels_gpu.py -q -x -p no:cacheprovider > gpurun_out/r2c2_temporal_tests.log 2>&1
tail -15 gpurun_out/r2c2_temporal_tests.log
timeout 400 python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_temporal_kernels_gpu.py > gpurun_out/r2c2_gpu_tests.log 2>&1
tail -8 gpurun_out/r2c2_gpu_tests.log
timeout 300 python tests/perf/temporal_microbench.py > gpurun_out/r2c2_temporal_microbench.log 2>&1
tail -70 gpurun_out/r2c2_temporal_microbench.log
timeout 400 python tests/perf/diagnose_legs.py > gpurun_out/r2c2_diagnose_legs.json 2> gpurun_out/r2c2_diagnose_legs.err
python - <<'P'
import json
r = json.load(open('gpurun_out/r2c2_diagnose_legs.json'))
for prec, d in r.items():
    print(prec)
    for k, v in d.items():
        print("  %-40s" % k, {a: ("%.2e" % b) for a, b in v.items()})
P
tail -3 gpurun_out/r2c2_diagnose_legs.err
timeout 200 python tests/perf/pipeline_breakdown.py 16 > gpurun_out/r2c2_breakdown.log 2>&1; tail -2 gpurun_out/r2c2_breakdown.log
timeout 200 python tests/perf/kernel_table.py 16 > /dev/null 2>&1; cp gpurun_out/kernel_table.txt gpurun_out/r2c2_kernel_table_T16.txt; head -40 gpurun_out/r2c2_kernel_table_T16.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c2_bench_n1.json 2> gpurun_out/r2c2_bench_n1.err
cat gpurun_out/r2c2_bench_n1.json; tail -3 gpurun_out/r2c2_bench_n1.err
