#!/bin/bash
# in-graph kernel times (single stream, CUPTI), launch list and one ncu --set full capture of the staged passes
mkdir -p gpurun_out
MOPOE_BRANCH_STREAMS=0 timeout 300 python bench.py --steps 3 --warmup 5 --no-cpu-baseline --profile-kernels > gpurun_out/r2p_in_graph_kernel_times.txt 2>&1
head -50 gpurun_out/r2p_in_graph_kernel_times.txt | cut -c1-110
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"staged_" -c 12 \
  -o gpurun_out/r2p_staged python tools/prof_ew.py > gpurun_out/r2p_ncu_staged.log 2>&1; echo "ncu exit $?"
