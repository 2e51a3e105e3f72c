#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r2g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2g_pytest.log
MOPOE_BENCH_SHAPES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2g_bench.log 2> gpurun_out/r2g_bench_shapes.log; echo "bench exit $?"
MOPOE_TC_TAPS=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2g_bench_notaps.log 2>&1
timeout 600 python bench.py --config 5-poe --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_cfg5poe.log 2>&1; echo "poe exit $?"
tail -n 6 gpurun_out/r2g_pytest.log
for f in r2g_bench r2g_bench_notaps r2g_bench_cfg5poe; do python - <<PY
import json
for l in open('gpurun_out/$f.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$f', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), round(d['roofline']['step_tensor_frac'],3))
PY
done
grep "N=16 \|im2col\|deconv3x3\|conv3x3" gpurun_out/r2g_bench_shapes.log
