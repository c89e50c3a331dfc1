#!/bin/bash
# round 2, GPU call 3: why are the temporal-stage kernels slow?  load-path probe + ncu of the kernels themselves
set -x
mkdir -p gpurun_out
timeout 120 tests/perf/probes/load_path_probe > gpurun_out/r2c3_load_path_probe.txt 2>&1
cat gpurun_out/r2c3_load_path_probe.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:'small_linear|flash_attn' -o gpurun_out/r2c3_ncu_temporal -f \
  python tests/perf/temporal_profile_target.py > gpurun_out/r2c3_ncu.log 2>&1
tail -5 gpurun_out/r2c3_ncu.log
ls -la gpurun_out | tail -5
