#!/bin/bash
# round 2, final scaling run on one 8-GPU box: the default path at N = 1, 2, 4, 8 back to back (like the driver's SCALE run)
mkdir -p gpurun_out
run() {  # n
  if [ "$1" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_scale_n1.json 2> gpurun_out/r2f_scale_n1.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2952$1 bench.py --gpus $1 \
      --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_scale_n$1.json 2> gpurun_out/r2f_scale_n$1.err
  fi
  python - <<P
import json
try:
    txt=open('gpurun_out/r2f_scale_n$1.json').read(); l=json.loads(txt[txt.index('{'):])
    print('N=$1', l['value'], l['ms_per_step'], 'e2e', l['e2e']['value'], l['e2e']['ms_per_step'], 'parity', l['parity_check']['timed_e2e_output_vs_eager_runner_rel_max_diff'], l['config'].get('temporal'))
except Exception as e:
    print('ERR N=$1', e); print(open('gpurun_out/r2f_scale_n$1.err').read()[-1500:])
P
}
run 1
run 2
run 4
run 8
