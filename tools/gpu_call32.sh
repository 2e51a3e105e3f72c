#!/bin/bash
mkdir -p gpurun_out
for v in 148 132 120 148b 140; do
  n=${v%b}
  if [ $n != 148 ]; then export MOPOE_GEMM_SMS=$n; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3e_bench_$v.log 2>&1
  unset MOPOE_GEMM_SMS
  python - <<PY
import json
for l in open('gpurun_out/r3e_bench_$v.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$v', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'clk', d['clocks']['sm_mhz'], 'gemm', round(d['roofline']['gemm_ms_per_step'],2), d['last_step']['total_loss'])
PY
done
