#!/bin/bash
# round 2, GPU call 12 (8 GPUs): scaling of the default path (round-robin temporal stage) at N = 4, 8; replicated at 8 for reference
set -x
mkdir -p gpurun_out
run() {  # n temporal
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2951$1 bench.py --gpus $1 \
    --steps 20 --warmup 5 --no-cpu-baseline --temporal $2 > gpurun_out/r2c12_bench_n$1_$2.json 2> gpurun_out/r2c12_bench_n$1_$2.err
  python - <<P
import json
try:
    txt=open('gpurun_out/r2c12_bench_n$1_$2.json').read(); l=json.loads(txt[txt.index('{'):])
    print('N=$1 $2', l['value'], l['ms_per_step'], 'e2e', l['e2e']['value'], l['e2e']['ms_per_step'], 'parity', l['parity_check']['timed_e2e_output_vs_eager_runner_rel_max_diff'])
except Exception as e:
    print('ERR N=$1 $2', e); print(open('gpurun_out/r2c12_bench_n$1_$2.err').read()[-2000:])
P
}
run 8 round_robin
run 4 round_robin
run 8 replicated
timeout 200 python -m pytest tests/test_zz_config5_clip_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -3
