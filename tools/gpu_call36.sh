#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q --maxfail=10 -k "text or wire or graphed" > gpurun_out/r3j_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3j_pytest.log
grep -E "^FAILED|passed|failed|exit|^E " gpurun_out/r3j_pytest.log | head -20
for v in onehot u8 u8nogather; do
  if [ $v = onehot ]; then A=""; else A="--text-wire uint8 --image-wire uint8"; fi
  if [ $v = u8nogather ]; then export MOPOE_TEXT_GATHER=0; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $A > gpurun_out/r3j_bench_$v.log 2>&1
  unset MOPOE_TEXT_GATHER
  python - <<PY
import json
for l in open('gpurun_out/r3j_bench_$v.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$v', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'h2d', d['e2e']['h2d_bytes_per_step'], 'clk', d['clocks']['sm_mhz'], d['last_step']['total_loss'])
PY
done
tail -n 5 gpurun_out/r3j_bench_u8.log | grep -v "^{" | tail -5
