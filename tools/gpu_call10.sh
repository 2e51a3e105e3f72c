#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -q --maxfail=10 > gpurun_out/r2i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2i_pytest.log
tail -n 4 gpurun_out/r2i_pytest.log
for v in default old u21 u43 u84; do
  echo "== $v"
  if [ $v = default ]; then timeout 300 python tools/prof_ew.py
  elif [ $v = old ]; then MOPOE_EW_ROWS=0 timeout 300 python tools/prof_ew.py
  else MOPOE_LIB_PATH=$PWD/tools/variants/lib_$v.so timeout 300 python tools/prof_ew.py; fi
done > gpurun_out/r2i_prof_ew.log 2>&1
cat gpurun_out/r2i_prof_ew.log
