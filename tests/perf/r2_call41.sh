#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; tail -2 gpurun_out/r2f_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_n1_reference.json 2> gpurun_out/r2f_bench_n1_reference.err
timeout 200 python tests/perf/kernel_table.py 16 > /dev/null 2>&1; cp gpurun_out/kernel_table.txt gpurun_out/r2f_kernel_table_T16.txt
python - <<'P'
import json
for f in ('r2f_bench_n1','r2f_bench_n1_reference'):
    l=json.load(open('gpurun_out/%s.json'%f))
    print(f, l.get('value'), l.get('ms_per_step'), 'e2e', (l.get('e2e') or {}).get('value'), 'parity', (l.get('parity_check') or {}).get('bit_identical'), 'cpu', l.get('cpu_baseline') and l['cpu_baseline'].get('value'), 'roof', (l.get('roofline') or {}).get('frac'), 'launches', l.get('gpu_launches'))
P
head -3 gpurun_out/r2f_kernel_table_T16.txt | cut -c1-120
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
