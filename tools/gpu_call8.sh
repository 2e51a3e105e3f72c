#!/bin/bash
# round-2 call 8: what bounds the BatchNorm / residual passes (plain timings, then ncu --set full of one launch each)
mkdir -p gpurun_out
timeout 300 python tools/prof_ew.py > gpurun_out/r2h_prof_ew.log 2>&1; echo "prof_ew exit $?"
cat gpurun_out/r2h_prof_ew.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"reduce_rows|bn_bwd_apply|bn_apply|combine" -c 14 \
  -o gpurun_out/r2h_ew python tools/prof_ew.py > gpurun_out/r2h_ncu_ew.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/*.ncu-rep
