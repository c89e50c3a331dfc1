#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tests/perf/chain_probe.py > gpurun_out/r2c5_chain_probe.log 2>&1
cat gpurun_out/r2c5_chain_probe.log | tail -80
