#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_elementwise.py -q --maxfail=50 > gpurun_out/r2m_ew_pytest.log 2>&1; echo "ew pytest exit $?" >> gpurun_out/r2m_ew_pytest.log
grep -E "^FAILED|passed|failed|exit" gpurun_out/r2m_ew_pytest.log | head -20
for v in k1 k8; do
MOPOE_LIB_PATH=$PWD/tools/variants/lib_$v.so timeout 600 python -m pytest tests/test_gpu_elementwise.py -q --maxfail=50 2>&1 | grep -E "^FAILED|passed|failed" | head -5
done
for v in default k1 k8; do
  echo "== $v"
  if [ $v = default ]; then timeout 300 python tools/prof_ew_shapes.py 2>&1 | head -8
  else MOPOE_LIB_PATH=$PWD/tools/variants/lib_$v.so timeout 300 python tools/prof_ew_shapes.py 2>&1 | head -8; fi
done | tee gpurun_out/r2m_shapes.txt
echo "== rows"; MOPOE_EW_STAGED=0 timeout 300 python tools/prof_ew_shapes.py 2>&1 | head -8
for v in default rows k1 k8; do
  if [ $v = default ]; then timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2m_bench_$v.log 2>&1
  elif [ $v = rows ]; then MOPOE_EW_STAGED=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2m_bench_$v.log 2>&1
  else MOPOE_LIB_PATH=$PWD/tools/variants/lib_$v.so timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2m_bench_$v.log 2>&1; fi
  python - <<PY
import json
for l in open('gpurun_out/r2m_bench_$v.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$v', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'clk', d['clocks']['sm_mhz'], {k:round(v['ms'],2) for k,v in d['roofline_hbm']['classes'].items() if k.startswith('bn') or k.startswith('comb')})
PY
done
