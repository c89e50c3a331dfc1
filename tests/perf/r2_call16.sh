#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_temporal_kernels_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 300 python tests/perf/temporal_microbench.py > gpurun_out/r2c16_temporal_microbench.log 2>&1
python - <<'P'
import json
r = json.load(open('gpurun_out/r2_temporal_microbench.json'))
for sec in ('stage_ms',):
    print(sec)
    for k, v in r[sec].items():
        print("   %-48s %s" % (k, v))
P
tail -5 gpurun_out/r2c16_temporal_microbench.log | cut -c1-300
