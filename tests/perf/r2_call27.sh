#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_linear_tc_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -5
timeout 200 python tests/perf/encoder_microbench.py 16 > gpurun_out/r2c27_encoder_microbench.json 2> gpurun_out/r2c27_encoder_microbench.err; cat gpurun_out/r2c27_encoder_microbench.json; tail -3 gpurun_out/r2c27_encoder_microbench.err

timeout 300 python -m pytest tests/test_mask_gemm_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 200 python tests/perf/mask_gemm_microbench.py > gpurun_out/r2c27_mask_gemm_microbench.log 2>&1; tail -5 gpurun_out/r2c27_mask_gemm_microbench.log
