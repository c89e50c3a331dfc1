#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_tc -c 4 -o gpurun_out/r2c26_ncu_linear_tc \
  python tests/perf/linear_tc_profile_target.py > gpurun_out/r2c26_ncu.log 2>&1
tail -3 gpurun_out/r2c26_ncu.log
ls -la gpurun_out/r2c26_ncu_linear_tc.ncu-rep
