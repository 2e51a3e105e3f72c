#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_gemm.py 256 2>&1 | tail -8
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"persist_kernel" -c 18 \
  -o gpurun_out/r3l_gemm python tools/prof_gemm.py 256 b1 > gpurun_out/r3l_ncu_gemm.log 2>&1; echo "ncu exit $?"
