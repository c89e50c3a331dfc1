#!/bin/bash
# N = 8 and N = 4 only (the GPU budget left at the end of round 2); same commands as r2_final_scale.sh
mkdir -p gpurun_out
for n in 8 4; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n \
    --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_scale_n$n.json 2> gpurun_out/r2f_scale_n$n.err
  python - <<P
import json
try:
    txt=open('gpurun_out/r2f_scale_n$n.json').read(); l=json.loads(txt[txt.index('{'):])
    print('N=$n', l['value'], l['ms_per_step'], 'e2e', l['e2e']['value'], 'parity', l['parity_check']['timed_e2e_output_vs_eager_runner_rel_max_diff'])
except Exception as e:
    print('ERR N=$n', e); print(open('gpurun_out/r2f_scale_n$n.err').read()[-1200:])
P
done
