#!/bin/bash
# Everything that was written after round 1's GPU budget ran out and still needs a number from a B200, in ONE gpurun call
# (1 GPU, ~12 min):   gpurun --timeout 900 -- 'bash tests/perf/pending_measurements.sh'
# Outputs land in gpurun_out/ (copy the summaries to profiles/ afterwards).
#   1. the full -m gpu suite with per-test durations (first run of test_zz_config5_gpu.py and of the late post-processing tests)
#   2. post-processing micro-benchmark: strip kernels vs the shared-memory tile variant (DVIS_VIS_MASKS_TILED), packed
#      variants, vps / vss kernels (never timed)
#   3. bench.py at N=1 (unchanged default path; 8-frame CPU baseline sample)
#   4. ncu launch list of one bench step of the FINAL pipeline (profiles/ only holds the first version's), launch-limited
#   5. ncu --set full of the post-processing kernels
# Two-GPU follow-up (separate call, --gpus 2): bench.py --temporal round_robin vs the replicated default:
#   gpurun --gpus 2 --timeout 600 -- 'for t in replicated round_robin; do python -m torch.distributed.run --nnodes=1 \
#     --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline \
#     --temporal $t > gpurun_out/bench_n2_$t.json; done'
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --durations=15 -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
tail -25 gpurun_out/gpu_tests.log
timeout 200 python tests/perf/postproc_microbench.py > gpurun_out/postproc_microbench.log 2>&1
tail -8 gpurun_out/postproc_microbench.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json
# the clip ending in fused VIS post-processing (different, smaller result: 18 MB of packed masks instead of 377 MB of logits)
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --d2h-stream > gpurun_out/bench_n1_d2hstream.json 2> gpurun_out/bench_n1_d2hstream.err
cat gpurun_out/bench_n1_d2hstream.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --d2h-stream --postprocess vis > gpurun_out/bench_n1_vis.json 2> gpurun_out/bench_n1_vis.err
cat gpurun_out/bench_n1_vis.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_final.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eager > gpurun_out/bench_under_ncu.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:vis_masks -c 6 -o gpurun_out/ncu_vis_masks \
  python tests/perf/postproc_microbench.py --quick > gpurun_out/ncu_vis_masks.log 2>&1
ls -la gpurun_out
