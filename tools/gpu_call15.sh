#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_elementwise.py -q --maxfail=50 > gpurun_out/r2n_ew_pytest.log 2>&1; echo "ew pytest exit $?" >> gpurun_out/r2n_ew_pytest.log
grep -E "^FAILED|passed|failed|exit|Error" gpurun_out/r2n_ew_pytest.log | head -20
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r2n_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2n_pytest.log
grep -E "^FAILED|passed|failed|exit" gpurun_out/r2n_pytest.log | head -30
for v in default nofuse; do
  if [ $v = default ]; then timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_bench_$v.log 2>&1
  else MOPOE_FUSE_NEXT_BN_STATS=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_bench_$v.log 2>&1; fi
  python - <<PY
import json
for l in open('gpurun_out/r2n_bench_$v.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$v', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'clk', d['clocks']['sm_mhz'], d['last_step'], {k:round(v['ms'],2) for k,v in d['roofline_hbm']['classes'].items() if k.startswith('bn') or k.startswith('comb')})
PY
done
