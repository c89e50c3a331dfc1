#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_zz_config5_clip_gpu.py -q -x -p no:cacheprovider > gpurun_out/r2c11_config5.log 2>&1
grep -B5 -A25 "Error" gpurun_out/r2c11_config5.log | cut -c1-300 | head -80
