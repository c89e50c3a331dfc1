#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; tail -2 gpurun_out/r2f_bench_n1.err
timeout 200 python tests/perf/kernel_table.py 16 > /dev/null 2>&1; cp gpurun_out/kernel_table.txt gpurun_out/r2f_kernel_table_T16.txt
python - <<'P'
import json
l=json.load(open('gpurun_out/r2f_bench_n1.json'))
print('final', l.get('value'), l.get('ms_per_step'), 'lat', l.get('latency_ms_per_clip'), 'e2e', (l.get('e2e') or {}).get('value'), 'parity', (l.get('parity_check') or {}).get('bit_identical'), 'cpu', l.get('cpu_baseline') and l['cpu_baseline'].get('value'), 'roof', (l.get('roofline') or {}).get('frac'), (l.get('roofline') or {}).get('us_per_launch'))
P
head -6 gpurun_out/r2f_kernel_table_T16.txt | cut -c1-130
