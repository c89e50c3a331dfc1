#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --postprocess vis > gpurun_out/r2c38_bench_vis.json 2> gpurun_out/r2c38_vis.err; tail -2 gpurun_out/r2c38_vis.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --eager > gpurun_out/r2c38_bench_eager.json 2> gpurun_out/r2c38_eager.err; tail -2 gpurun_out/r2c38_eager.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c38_bench_n1.json 2> gpurun_out/r2c38_n1.err
python - <<'P'
import json
for f in ('r2c38_bench_vis','r2c38_bench_eager','r2c38_bench_n1'):
    try:
        l=json.load(open('gpurun_out/%s.json'%f))
        print(f, l.get('value'), l.get('ms_per_step'), 'e2e', (l.get('e2e') or {}).get('value'), (l.get('e2e') or {}).get('d2h_bytes_per_step'), 'parity', (l.get('parity_check') or {}).get('bit_identical'))
    except Exception as e:
        print(f, 'ERR', e)
P
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
