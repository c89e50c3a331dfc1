#!/bin/bash
# round 2, final single-GPU measurements: default bench (with cpu_baseline), reference arm, fp32 line, configs 2 / 3,
# kernel table + ncu launch list of one eager step of the FINAL pipeline
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; tail -2 gpurun_out/r2f_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_n1_reference.json 2> gpurun_out/r2f_bench_n1_reference.err; tail -2 gpurun_out/r2f_bench_n1_reference.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --precision fp32 > gpurun_out/r2f_bench_n1_fp32.json 2> gpurun_out/r2f_fp32.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config 2 > gpurun_out/r2f_bench_config2.json 2> gpurun_out/r2f_c2.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config 3 > gpurun_out/r2f_bench_config3.json 2> gpurun_out/r2f_c3.err
timeout 200 python tests/perf/kernel_table.py 16 > /dev/null 2>&1; cp gpurun_out/kernel_table.txt gpurun_out/r2f_kernel_table_T16.txt
timeout 200 python tests/perf/pipeline_breakdown.py 16 > gpurun_out/r2f_breakdown.log 2>&1; tail -1 gpurun_out/r2f_breakdown.log
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r2f_launches_final.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eager > gpurun_out/r2f_bench_under_ncu.log 2>&1
python - <<'P'
import json
for f in ('r2f_bench_n1','r2f_bench_n1_reference','r2f_bench_n1_fp32','r2f_bench_config2','r2f_bench_config3'):
    try:
        l=json.load(open('gpurun_out/%s.json'%f))
        print(f, l.get('value'), l.get('ms_per_step'), 'e2e', (l.get('e2e') or {}).get('value'), 'parity', (l.get('parity_check') or {}).get('bit_identical'), 'cpu', l.get('cpu_baseline') and l['cpu_baseline'].get('value'))
    except Exception as e:
        print(f, 'ERR', e)
P
head -12 gpurun_out/r2f_kernel_table_T16.txt | cut -c1-150
wc -l gpurun_out/r2f_launches_final.csv
