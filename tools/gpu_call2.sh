#!/bin/bash
# round-2 GPU call 2: TMA-store epilogue + fused BatchNorm statistics
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py -q --maxfail=10 -x > gpurun_out/r2b_gemm.log 2>&1; echo "gemm exit $?" >> gpurun_out/r2b_gemm.log
if grep -q "gemm exit 0" gpurun_out/r2b_gemm.log; then
  timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 --deselect tests/test_gpu_gemm.py > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest.log
  MOPOE_BENCH_SHAPES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench.log 2> gpurun_out/r2b_bench_shapes.log; echo "bench exit $?" >> gpurun_out/r2b_bench_shapes.log
  MOPOE_BRANCH_STREAMS=0 timeout 300 python bench.py --steps 3 --warmup 5 --no-cpu-baseline --profile-kernels > gpurun_out/r2b_in_graph_kernel_times.txt 2>&1
  MOPOE_GEMM_TMA_EPI=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_noepi.log 2>&1
  MOPOE_FUSE_BN_STATS=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_nostats.log 2>&1
fi
tail -n 15 gpurun_out/r2b_gemm.log
tail -n 8 gpurun_out/r2b_pytest.log
tail -c 600 gpurun_out/r2b_bench.log
