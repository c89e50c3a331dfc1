#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_msda_gpu.py tests/test_msda_backward_gpu.py tests/test_dropin_reference.py tests/test_abi.py -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 300 python tests/perf/msda_microbench.py > gpurun_out/r2c39_msda_microbench.log 2>&1; tail -12 gpurun_out/r2c39_msda_microbench.log | cut -c1-250
