#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_fused_epi.py 256 > gpurun_out/r4i_prof.log 2>&1; echo "prof exit $?"; cat gpurun_out/r4i_prof.log
PROF_ITERS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"persist_kernel" -c 12 \
  -o gpurun_out/r4i_fused python tools/prof_fused_epi.py 256 > gpurun_out/r4i_ncu.log 2>&1; echo "ncu exit $?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r4i_bench.log 2> gpurun_out/r4i_bench.err; echo "bench exit $?"; tail -c 1500 gpurun_out/r4i_bench.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r4i_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/r4i_ncu_launches.log 2>&1; echo "launch list exit $?"
ls -la gpurun_out/r4i_*
