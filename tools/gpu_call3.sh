#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py -q --maxfail=10 -x > gpurun_out/r2c_gemm.log 2>&1; echo "gemm exit $?" >> gpurun_out/r2c_gemm.log
if grep -q "gemm exit 0" gpurun_out/r2c_gemm.log; then
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -q --maxfail=10 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest.log
  MOPOE_BENCH_SHAPES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench.log 2> gpurun_out/r2c_bench_shapes.log; echo "bench exit $?" >> gpurun_out/r2c_bench_shapes.log
  MOPOE_FUSE_BN_STATS=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_nostats.log 2>&1
  MOPOE_GEMM_BM256=1 timeout 300 python -m pytest tests/test_gpu_gemm.py -q -x > gpurun_out/r2c_gemm_bm256.log 2>&1; echo "gemm exit $?" >> gpurun_out/r2c_gemm_bm256.log
  MOPOE_GEMM_BM256=1 MOPOE_BENCH_SHAPES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_bm256.log 2> gpurun_out/r2c_bench_bm256_shapes.log
fi
tail -n 5 gpurun_out/r2c_gemm.log gpurun_out/r2c_pytest.log gpurun_out/r2c_gemm_bm256.log
for f in r2c_bench r2c_bench_nostats r2c_bench_bm256; do python - <<PY
import json
for l in open('gpurun_out/$f.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$f', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'gemm ms', round(d['roofline']['gemm_ms_per_step'],2), {k:round(v['ms'],2) for k,v in d['roofline']['by_kind'].items()})
PY
done
