#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py -q --maxfail=10 -k "tc or bf16" > gpurun_out/r3m_gemm.log 2>&1; echo "gemm exit $?" >> gpurun_out/r3m_gemm.log
grep -E "^FAILED|passed|failed|exit|^E " gpurun_out/r3m_gemm.log | head -10
if grep -q "gemm exit 0" gpurun_out/r3m_gemm.log; then
for v in default bn128 default2 bn128b; do
  if [ $v = bn128 -o $v = bn128b ]; then export MOPOE_LIB_PATH=$PWD/tools/variants/lib_bn128.so; fi
  MOPOE_BENCH_SHAPES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3m_bench_$v.log 2> gpurun_out/r3m_shapes_$v.log
  unset MOPOE_LIB_PATH
  python - <<PY
import json
for l in open('gpurun_out/r3m_bench_$v.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$v', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'clk', d['clocks']['sm_mhz'], 'gemm', round(d['roofline']['gemm_ms_per_step'],2), d['last_step']['total_loss'])
PY
done
grep "N=640" gpurun_out/r3m_shapes_default.log | grep -v "^wg" | head -8; echo; grep "N=640" gpurun_out/r3m_shapes_bn128.log | grep -v "^wg" | head -8
fi
