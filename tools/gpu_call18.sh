#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r2o_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2o_pytest.log
grep -E "^FAILED|passed|failed|exit" gpurun_out/r2o_pytest.log | head -30
echo "== reverse"; timeout 300 python tools/prof_ew.py 2>&1 | grep -v torch
echo "== forward"; MOPOE_ST_REVERSE=0 timeout 300 python tools/prof_ew.py 2>&1 | grep -v torch
for v in default forward; do
  if [ $v = default ]; then timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2o_bench_$v.log 2>&1
  else MOPOE_ST_REVERSE=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2o_bench_$v.log 2>&1; fi
  python - <<PY
import json
for l in open('gpurun_out/r2o_bench_$v.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$v', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'clk', d['clocks']['sm_mhz'], 'gemm', round(d['roofline']['gemm_ms_per_step'],2), {k:round(v['ms'],2) for k,v in d['roofline_hbm']['classes'].items() if k.startswith('bn') or k.startswith('comb')})
PY
done
