#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py -q --maxfail=20 -k "backward_sums or residual_combine or fused_batchnorm" > gpurun_out/r4f_gemm.log 2>&1; echo "gemm exit $?" >> gpurun_out/r4f_gemm.log
grep -E "^FAILED|passed|failed|exit|^E " gpurun_out/r4f_gemm.log | head -40
for v in new prev new2 prev2; do
  if [ $v = prev -o $v = prev2 ]; then export MOPOE_LIB_PATH=$PWD/tools/variants/lib_prev.so; fi
  MOPOE_BENCH_SHAPES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r4f_bench_$v.log 2> gpurun_out/r4f_shapes_$v.log
  echo "bench $v exit $?"
  unset MOPOE_LIB_PATH
  python - <<PY
import json
for l in open('gpurun_out/r4f_bench_$v.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$v', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'clk', d['clocks']['sm_mhz'], 'gemm', round(d['roofline']['gemm_ms_per_step'],2), d['last_step']['total_loss'])
PY
done
grep -E "bnb|res" gpurun_out/r4f_shapes_new.log | head -12
echo; grep -E "bnb|res" gpurun_out/r4f_shapes_prev.log | head -12
