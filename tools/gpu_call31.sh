#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_elementwise.py -q --maxfail=20 > gpurun_out/r3d_ew.log 2>&1; echo "ew exit $?" >> gpurun_out/r3d_ew.log
grep -E "^FAILED|passed|failed|exit|^E " gpurun_out/r3d_ew.log | head -20
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 --deselect tests/test_gpu_elementwise.py > gpurun_out/r3d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3d_pytest.log
grep -E "^FAILED|passed|failed|exit" gpurun_out/r3d_pytest.log | head -30
for v in default norecomp default2 norecomp2; do
  if [ $v = norecomp -o $v = norecomp2 ]; then export MOPOE_GATE_RECOMPUTE=0; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3d_bench_$v.log 2>&1
  unset MOPOE_GATE_RECOMPUTE
  python - <<PY
import json
for l in open('gpurun_out/r3d_bench_$v.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$v', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'clk', d['clocks']['sm_mhz'], {k:round(v['ms'],2) for k,v in d['roofline_hbm']['classes'].items() if k.startswith('bn_bwd')}, d['last_step']['total_loss'])
PY
done
