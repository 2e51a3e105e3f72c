#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py -q --maxfail=10 -k "tc or bf16" > gpurun_out/r2y_gemm.log 2>&1; echo "gemm exit $?" >> gpurun_out/r2y_gemm.log
grep -E "^FAILED|passed|failed|exit|Error|error" gpurun_out/r2y_gemm.log | head -20
if grep -q "gemm exit 0" gpurun_out/r2y_gemm.log; then
for v in default preelect default2 preelect2; do
  if [ $v = preelect -o $v = preelect2 ]; then export MOPOE_LIB_PATH=$PWD/tools/variants/lib_preelect.so; fi
  MOPOE_BENCH_SHAPES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2y_bench_$v.log 2> gpurun_out/r2y_shapes_$v.log
  unset MOPOE_LIB_PATH
  python - <<PY
import json
for l in open('gpurun_out/r2y_bench_$v.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$v', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'clk', d['clocks']['sm_mhz'], 'gemm', round(d['roofline']['gemm_ms_per_step'],2), {k:round(v['ms'],2) for k,v in d['roofline']['by_kind'].items()}, d['last_step']['total_loss'])
PY
done
head -14 gpurun_out/r2y_shapes_default.log; echo; head -14 gpurun_out/r2y_shapes_preelect.log
fi
