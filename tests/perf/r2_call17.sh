#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tests/perf/tracker_kernel_times.py > gpurun_out/r2c17_tracker_kernel_times.txt 2>&1
cat gpurun_out/r2c17_tracker_kernel_times.txt | cut -c1-220
