#!/bin/bash
# round-2 GPU call 1: full GPU test-suite, smoke, bf16 per-tensor parity report, bench with the shape table, in-graph kernel
# times, and the ncu --set full capture of the likelihood / fusion kernels
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r2_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest.log
timeout 400 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r2_smoke.log
timeout 1200 python tools/bf16_parity.py > gpurun_out/r2_bf16_parity.log 2>&1; echo "exit $?" >> gpurun_out/r2_bf16_parity.log
MOPOE_BENCH_SHAPES=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.log 2> gpurun_out/r2_bench_shapes.log; echo "bench exit $?" >> gpurun_out/r2_bench_shapes.log
MOPOE_BRANCH_STREAMS=0 timeout 300 python bench.py --steps 3 --warmup 5 --no-cpu-baseline --profile-kernels > gpurun_out/r2_in_graph_kernel_times.txt 2>&1
timeout 200 python tools/prof_nll.py > gpurun_out/r2_prof_nll.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"categorical|laplace|fusion" -c 16 -f -o gpurun_out/r2_nll python tools/prof_nll.py > gpurun_out/r2_ncu_nll.log 2>&1
tail -5 gpurun_out/r2_pytest.log gpurun_out/r2_smoke.log gpurun_out/r2_bench.log gpurun_out/r2_prof_nll.log
tail -30 gpurun_out/r2_bf16_parity.log
