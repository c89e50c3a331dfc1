#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_mask_gemm_gpu.py tests/test_temporal_kernels_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -6
timeout 300 python tests/perf/mask_gemm_microbench.py > gpurun_out/r2c31_mask_gemm_microbench.log 2>&1; grep -v "fp32_nchw" gpurun_out/r2c31_mask_gemm_microbench.log | cut -c1-260
cp gpurun_out/mask_gemm_microbench.json gpurun_out/r2c31_mask_gemm_microbench.json
