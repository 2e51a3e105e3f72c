#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"oneshot|combine_kernel" -c 8 \
  -o gpurun_out/r2h_ew2 python tools/prof_ew.py > gpurun_out/r2h_ncu_ew2.log 2>&1; echo "ncu exit $?"
