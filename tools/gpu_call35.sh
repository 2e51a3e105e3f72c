#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -q --maxfail=10 > gpurun_out/r3i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3i_pytest.log
grep -E "^FAILED|passed|failed|exit|^E " gpurun_out/r3i_pytest.log | head -20
for v in default nov5 default2 nov5b; do
  if [ $v = nov5 -o $v = nov5b ]; then export MOPOE_WGRAD_FINISH_V5=0; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3i_bench_$v.log 2>&1
  unset MOPOE_WGRAD_FINISH_V5
  python - <<PY
import json
for l in open('gpurun_out/r3i_bench_$v.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$v', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'clk', d['clocks']['sm_mhz'], {k:round(v['ms'],2) for k,v in d['roofline']['by_kind'].items()}, d['last_step']['total_loss'])
PY
done
