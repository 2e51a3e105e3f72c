#!/bin/bash
mkdir -p gpurun_out
for v in default pair128 default2 pair128b; do
  if [ $v = pair128 -o $v = pair128b ]; then export MOPOE_GEMM_PAIR_MIN_BN=64; fi
  MOPOE_BENCH_SHAPES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2z_bench_$v.log 2> gpurun_out/r2z_shapes_$v.log
  unset MOPOE_GEMM_PAIR_MIN_BN
  python - <<PY
import json
for l in open('gpurun_out/r2z_bench_$v.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$v', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'clk', d['clocks']['sm_mhz'], 'gemm', round(d['roofline']['gemm_ms_per_step'],2), {k:round(v['ms'],2) for k,v in d['roofline']['by_kind'].items()}, d['last_step']['total_loss'])
PY
done
grep "N=128 " gpurun_out/r2z_shapes_default.log | head -8; echo; grep "N=128 " gpurun_out/r2z_shapes_pair128.log | head -8
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r2z_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2z_pytest.log
grep -E "^FAILED|passed|failed|exit" gpurun_out/r2z_pytest.log | head -30
