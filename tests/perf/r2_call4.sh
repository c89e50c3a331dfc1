#!/bin/bash
# round 2, GPU call 4: re-tuned temporal kernels (8 warps, hoisted addressing, deeper ring), predictor on flash attention + bit masks
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_temporal_kernels_gpu.py -q -x -p no:cacheprovider > gpurun_out/r2c4_temporal_tests.log 2>&1
tail -15 gpurun_out/r2c4_temporal_tests.log
timeout 400 python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_temporal_kernels_gpu.py > gpurun_out/r2c4_gpu_tests.log 2>&1
tail -8 gpurun_out/r2c4_gpu_tests.log
timeout 300 python tests/perf/temporal_microbench.py > gpurun_out/r2c4_temporal_microbench.log 2>&1
tail -120 gpurun_out/r2c4_temporal_microbench.log
timeout 200 python tests/perf/pipeline_breakdown.py 16 > gpurun_out/r2c4_breakdown.log 2>&1; tail -2 gpurun_out/r2c4_breakdown.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:'small_linear|flash_attn' -o gpurun_out/r2c4_ncu_temporal -f \
  python tests/perf/temporal_profile_target.py > gpurun_out/r2c4_ncu.log 2>&1
tail -3 gpurun_out/r2c4_ncu.log
