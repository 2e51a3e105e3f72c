#!/bin/bash
mkdir -p gpurun_out
for v in sort nosort sort2 nosort2; do
  if [ $v = nosort -o $v = nosort2 ]; then export MOPOE_PACK_SORT=0; fi
  MOPOE_BENCH_SHAPES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r4h_bench_$v.log 2> gpurun_out/r4h_shapes_$v.log
  echo "bench $v exit $?"
  unset MOPOE_PACK_SORT
  python - <<PY
import json
for l in open('gpurun_out/r4h_bench_$v.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$v', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'clk', d['clocks']['sm_mhz'], 'gemm', round(d['roofline']['gemm_ms_per_step'],2), d['last_step']['total_loss'])
PY
  grep "pack" gpurun_out/r4h_shapes_$v.log | head -3
done
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r4h_suite.log 2>&1; echo "suite exit $?" >> gpurun_out/r4h_suite.log
grep -E "^FAILED|passed|failed|exit|^E " gpurun_out/r4h_suite.log | head -20
