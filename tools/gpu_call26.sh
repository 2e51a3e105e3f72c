#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py -q --maxfail=10 -k "tc or bf16" > gpurun_out/r2x_gemm.log 2>&1; echo "gemm exit $?" >> gpurun_out/r2x_gemm.log
grep -E "^FAILED|passed|failed|exit|Error|error" gpurun_out/r2x_gemm.log | head -20
if grep -q "gemm exit 0" gpurun_out/r2x_gemm.log; then
for v in default nowgpair default2 nowgpair2; do
  if [ $v = nowgpair -o $v = nowgpair2 ]; then export MOPOE_WGRAD_PAIR=0; fi
  MOPOE_BENCH_SHAPES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2x_bench_$v.log 2> gpurun_out/r2x_shapes_$v.log
  unset MOPOE_WGRAD_PAIR
  python - <<PY
import json
for l in open('gpurun_out/r2x_bench_$v.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$v', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'clk', d['clocks']['sm_mhz'], 'gemm', round(d['roofline']['gemm_ms_per_step'],2), {k:round(v['ms'],2) for k,v in d['roofline']['by_kind'].items()}, d['last_step']['total_loss'])
PY
done
grep "^wg" gpurun_out/r2x_shapes_default.log | head -12; echo; grep "^wg" gpurun_out/r2x_shapes_nowgpair.log | head -12
fi
