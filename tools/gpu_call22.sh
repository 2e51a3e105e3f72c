#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py -q --maxfail=10 -x > gpurun_out/r2s_gemm.log 2>&1; echo "gemm exit $?" >> gpurun_out/r2s_gemm.log
tail -n 3 gpurun_out/r2s_gemm.log
if grep -q "gemm exit 0" gpurun_out/r2s_gemm.log; then
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 --deselect tests/test_gpu_gemm.py > gpurun_out/r2s_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2s_pytest.log
grep -E "^FAILED|passed|failed|exit" gpurun_out/r2s_pytest.log | head -30
for v in default oldgemm default2 oldgemm2; do
  if [ $v = oldgemm -o $v = oldgemm2 ]; then export MOPOE_LIB_PATH=$PWD/tools/variants/lib_oldgemm.so; fi
  MOPOE_BENCH_SHAPES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2s_bench_$v.log 2> gpurun_out/r2s_shapes_$v.log
  unset MOPOE_LIB_PATH
  python - <<PY
import json
for l in open('gpurun_out/r2s_bench_$v.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$v', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'clk', d['clocks']['sm_mhz'], 'gemm', round(d['roofline']['gemm_ms_per_step'],2), {k:round(v['ms'],2) for k,v in d['roofline']['by_kind'].items()}, d['last_step']['total_loss'])
PY
done
grep "+bn" gpurun_out/r2s_shapes_default.log | head -12
echo; grep "+bn" gpurun_out/r2s_shapes_oldgemm.log | head -12
fi
