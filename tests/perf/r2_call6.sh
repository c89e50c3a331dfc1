#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_temporal_kernels_gpu.py -q -x -p no:cacheprovider > gpurun_out/r2c6_temporal_tests.log 2>&1
tail -5 gpurun_out/r2c6_temporal_tests.log
timeout 300 python tests/perf/chain_probe.py > gpurun_out/r2c6_chain_probe.log 2>&1
tail -75 gpurun_out/r2c6_chain_probe.log
timeout 300 python tests/perf/temporal_microbench.py > gpurun_out/r2c6_temporal_microbench.log 2>&1
python - <<'P'
import json
r = json.load(open('gpurun_out/r2_temporal_microbench.json'))
for sec, d in r.items():
    print(sec)
    for k, v in d.items():
        print("   %-48s %s" % (k, v))
P
