#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py -q --maxfail=10 -k "tc or bf16 or split" > gpurun_out/r3f_gemm.log 2>&1; echo "gemm exit $?" >> gpurun_out/r3f_gemm.log
grep -E "^FAILED|passed|failed|exit|^E " gpurun_out/r3f_gemm.log | head -20
if grep -q "gemm exit 0" gpurun_out/r3f_gemm.log; then
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 --deselect tests/test_gpu_gemm.py > gpurun_out/r3f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3f_pytest.log
grep -E "^FAILED|passed|failed|exit" gpurun_out/r3f_pytest.log | head -30
for v in default nosplit default2 nosplit2; do
  if [ $v = nosplit -o $v = nosplit2 ]; then export MOPOE_GEMM_SPLITK=0; fi
  MOPOE_BENCH_SHAPES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3f_bench_$v.log 2> gpurun_out/r3f_shapes_$v.log
  unset MOPOE_GEMM_SPLITK
  python - <<PY
import json
for l in open('gpurun_out/r3f_bench_$v.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$v', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'clk', d['clocks']['sm_mhz'], 'gemm', round(d['roofline']['gemm_ms_per_step'],2), d['last_step']['total_loss'])
PY
done
grep "^sk" gpurun_out/r3f_shapes_default.log | head; echo; grep "M=256x1x1 N=640 K=4x\|M=256x1x4 N=640\|M=256x1x1 N=640 K=1x2560" gpurun_out/r3f_shapes_nosplit.log | head
fi
