#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r2q_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2q_pytest.log
grep -E "^FAILED|passed|failed|exit" gpurun_out/r2q_pytest.log | head -30
for v in default nomask nofin default2; do
  if [ $v = nomask ]; then export MOPOE_BATCHED_MASKS=0; elif [ $v = nofin ]; then export MOPOE_WGRAD_FINISH_V4=0; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2q_bench_$v.log 2>&1
  unset MOPOE_BATCHED_MASKS MOPOE_WGRAD_FINISH_V4
  python - <<PY
import json
for l in open('gpurun_out/r2q_bench_$v.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$v', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'clk', d['clocks']['sm_mhz'], 'gemm', round(d['roofline']['gemm_ms_per_step'],2), {k:round(v['ms'],2) for k,v in d['roofline']['by_kind'].items()}, d['last_step']['total_loss'])
PY
done
