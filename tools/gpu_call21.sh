#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r2r_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2r_pytest.log
grep -E "^FAILED|passed|failed|exit" gpurun_out/r2r_pytest.log | head -30
for v in default nows default2 nows2; do
  if [ $v = nows -o $v = nows2 ]; then export MOPOE_WGRAD_STREAMS=0; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2r_bench_$v.log 2>&1
  unset MOPOE_WGRAD_STREAMS
  python - <<PY
import json
for l in open('gpurun_out/r2r_bench_$v.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$v', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'clk', d['clocks']['sm_mhz'], d['last_step']['total_loss'])
PY
done
tail -3 gpurun_out/r2r_bench_default.log | cut -c1-300
