#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_temporal_kernels_gpu.py tests/test_modules_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 300 python tests/perf/chain_probe.py > gpurun_out/r2c13_chain_probe.log 2>&1
tail -62 gpurun_out/r2c13_chain_probe.log
timeout 300 python tests/perf/temporal_microbench.py > gpurun_out/r2c13_temporal_microbench.log 2>&1
python - <<'P'
import json
r = json.load(open('gpurun_out/r2_temporal_microbench.json'))
for sec in ('linear_us', 'stage_ms'):
    print(sec)
    for k, v in r[sec].items():
        print("   %-48s %s" % (k, v))
P
timeout 200 python tests/perf/pipeline_breakdown.py 16 > gpurun_out/r2c13_breakdown.log 2>&1; tail -1 gpurun_out/r2c13_breakdown.log
