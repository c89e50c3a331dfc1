#!/bin/bash
# round 2, GPU call 10 (2 GPUs): the frame-sharded bench with the round-robin temporal stage (default at N > 1) and replicated
set -x
mkdir -p gpurun_out
for t in round_robin replicated; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 \
    --steps 20 --warmup 5 --no-cpu-baseline --temporal $t > gpurun_out/r2c10_bench_n2_$t.json 2> gpurun_out/r2c10_bench_n2_$t.err
  python - <<P
import json
try:
    l=json.load(open('gpurun_out/r2c10_bench_n2_$t.json'))
    print('$t', l['value'], l['ms_per_step'], 'e2e', l['e2e']['value'], l['e2e']['ms_per_step'], 'parity', l['parity_check'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r2c10_bench_n2_$t.err').read()[-2500:])
P
done
timeout 300 python -m pytest tests/test_zz_config5_clip_gpu.py tests/test_msda_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c10_bench_n1.json 2> gpurun_out/r2c10_bench_n1.err
python - <<'P'
import json
l=json.load(open('gpurun_out/r2c10_bench_n1.json')); r=l['roofline']
print('n1', l['value'], l['ms_per_step'], 'e2e', l['e2e']['value'], 'parity', l['parity_check']['bit_identical'], 'msda us', r['us_per_launch'], 'frac', r['frac'])
P
