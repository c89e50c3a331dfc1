#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tests/perf/chain_probe.py > gpurun_out/r2c15_chain_probe.log 2>&1
grep -A12 "chain_us_per_step" gpurun_out/r2c15_chain_probe.log | head -40; grep "linear x32\|dvis_flash_attn\|torch_sdpa" gpurun_out/r2c15_chain_probe.log
timeout 300 python tests/perf/temporal_microbench.py > gpurun_out/r2c15_temporal_microbench.log 2>&1
python - <<'P'
import json
r = json.load(open('gpurun_out/r2_temporal_microbench.json'))
for sec in ('attention_us', 'linear_us', 'stage_ms'):
    print(sec)
    for k, v in r[sec].items():
        print("   %-48s %s" % (k, v))
P
for i in 1 2; do timeout 200 python tests/perf/pipeline_breakdown.py 16 > gpurun_out/r2c15_breakdown.log 2>&1; tail -1 gpurun_out/r2c15_breakdown.log; done
timeout 200 python tests/perf/kernel_table.py 16 > /dev/null 2>&1; cp gpurun_out/kernel_table.txt gpurun_out/r2c15_kernel_table_T16.txt; head -24 gpurun_out/r2c15_kernel_table_T16.txt | cut -c1-150
