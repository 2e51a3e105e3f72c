#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_elementwise.py -q --maxfail=50 > gpurun_out/r2j_ew_pytest.log 2>&1; echo "ew pytest exit $?" >> gpurun_out/r2j_ew_pytest.log
grep -E "^FAILED|passed|failed" gpurun_out/r2j_ew_pytest.log | head -40
MOPOE_EW_STAGED=0 timeout 600 python -m pytest tests/test_gpu_elementwise.py -q --maxfail=5 -x > gpurun_out/r2j_ew_pytest_rows.log 2>&1; echo "ew(rows) pytest exit $?" >> gpurun_out/r2j_ew_pytest_rows.log
tail -n 5 gpurun_out/r2j_ew_pytest_rows.log
for v in staged rows; do
  echo "== $v"
  if [ $v = staged ]; then timeout 300 python tools/prof_ew.py; else MOPOE_EW_STAGED=0 timeout 300 python tools/prof_ew.py; fi
done > gpurun_out/r2j_prof_ew.log 2>&1
cat gpurun_out/r2j_prof_ew.log
if grep -q "ew pytest exit 0" gpurun_out/r2j_ew_pytest.log; then
  timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -q --maxfail=10 > gpurun_out/r2j_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2j_pytest.log
  tail -n 4 gpurun_out/r2j_pytest.log
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2j_bench.log 2>&1; echo "bench exit $?"
  MOPOE_EW_STAGED=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2j_bench_rows.log 2>&1
  for f in r2j_bench r2j_bench_rows; do python - <<PY
import json
for l in open('gpurun_out/$f.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$f', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), round(d['roofline']['step_tensor_frac'],3), {k:(round(v['ms'],2),round(v['frac'],2)) for k,v in d['roofline_hbm']['classes'].items()})
PY
  done
fi
