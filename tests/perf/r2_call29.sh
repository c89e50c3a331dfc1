#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2 3 4 5; do DVIS_MSDA_HM_VARIANT=$v timeout 100 python tests/perf/msda_hm_variants.py 16; done 2>&1 | tee gpurun_out/r2c29_msda_hm_variants.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -c 1 --launch-skip 3 -o gpurun_out/r2c29_ncu_msda_hm \
  python tests/perf/msda_hm_variants.py 8 > gpurun_out/r2c29_ncu.log 2>&1
tail -2 gpurun_out/r2c29_ncu.log
