#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none -k regex:"mask_gemm|linear_tc" --launch-skip 6 -c 6 -o gpurun_out/r2c45_ncu_tensor_kernels \
  python tests/perf/tensor_kernels_profile_target.py > gpurun_out/r2c45_ncu.log 2>&1
tail -2 gpurun_out/r2c45_ncu.log
ncu -i gpurun_out/r2c45_ncu_tensor_kernels.ncu-rep --page raw --csv > gpurun_out/r2c45_ncu_tensor_kernels_raw.csv 2>/dev/null
wc -c gpurun_out/r2c45_ncu_tensor_kernels_raw.csv
