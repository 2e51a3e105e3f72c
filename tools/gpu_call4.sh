#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_kernels.py -q --maxfail=10 > gpurun_out/r2d_kernels.log 2>&1; echo "kernels exit $?" >> gpurun_out/r2d_kernels.log
if grep -q "kernels exit 0" gpurun_out/r2d_kernels.log; then
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bf16.py tests/test_gpu_eval.py -q --maxfail=10 > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_pytest.log
  MOPOE_BENCH_SHAPES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench.log 2> gpurun_out/r2d_bench_shapes.log; echo "bench exit $?" >> gpurun_out/r2d_bench_shapes.log
  MOPOE_BRANCH_STREAMS=0 timeout 300 python bench.py --steps 3 --warmup 5 --no-cpu-baseline --profile-kernels > gpurun_out/r2d_in_graph_kernel_times.txt 2>&1
fi
tail -n 6 gpurun_out/r2d_kernels.log gpurun_out/r2d_pytest.log
python - <<PY
import json
for l in open('gpurun_out/r2d_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'gemm ms', round(d['roofline']['gemm_ms_per_step'],2), {k:round(v['ms'],2) for k,v in d['roofline']['by_kind'].items()})
PY
grep -v "^x\|^wg" gpurun_out/r2d_bench_shapes.log | head -12
