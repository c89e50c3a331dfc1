#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  Run the emulated kernels under a sanitizer:
#   tests/simt/run_sanitized.sh address [pytest args]   every load / store an emulated CUDA thread makes (global or "shared"
#                                                        memory) is bounds-checked -- stand-in for compute-sanitizer memcheck
#   tests/simt/run_sanitized.sh thread  [pytest args]   happens-before race detection between the emulated threads, with
#                                                        __syncthreads() / warp collectives / atomics as the only ordering --
#                                                        stand-in for compute-sanitizer racecheck (shared AND global memory).
#                                                        Reports go to /tmp/simt_tsan.<pid>; no file = no race.
set -e
mode=${1:-address}; shift || true
cd "$(dirname "$0")/../.."
export SIMT_SANITIZE=$mode
if [ "$mode" = thread ]; then
  rm -f /tmp/simt_tsan.*
  export TSAN_OPTIONS="halt_on_error=0:exitcode=0:log_path=/tmp/simt_tsan" OMP_NUM_THREADS=1
  LD_PRELOAD=$(gcc -print-file-name=libtsan.so) python -m pytest tests/test_simt_kernels.py tests/test_simt_msda_and_norms.py tests/test_simt_nonfinite.py -q "$@"
  if ls /tmp/simt_tsan.* >/dev/null 2>&1; then grep -h SUMMARY /tmp/simt_tsan.* | sort | uniq -c; exit 1; fi
  echo "ThreadSanitizer: no data race reported"
else
  export ASAN_OPTIONS=detect_leaks=0
  LD_PRELOAD=$(gcc -print-file-name=libasan.so) python -m pytest tests/test_simt_kernels.py tests/test_simt_msda_and_norms.py tests/test_simt_nonfinite.py -q "$@"
fi
