#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  Run the emulated kernels under AddressSanitizer: every load / store an emulated CUDA thread
# makes (global or "shared" memory) is bounds-checked against the host allocations -- the CPU stand-in for
# `compute-sanitizer --tool memcheck`.  Usage: tests/simt/run_sanitized.sh [pytest args]
set -e
cd "$(dirname "$0")/../.."
export SIMT_SANITIZE=address
export ASAN_OPTIONS=detect_leaks=0
LD_PRELOAD=$(gcc -print-file-name=libasan.so) python -m pytest tests/test_simt_kernels.py tests/test_simt_msda_and_norms.py -q "$@"
