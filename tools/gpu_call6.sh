#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.log 2> gpurun_out/r2f_bench.err; echo "bench exit $?"
MOPOE_DP_BUCKETS=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench_nooverlap.log 2>&1
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2f_bench_ref.log 2>&1; echo "ref exit $?"
tail -n 4 gpurun_out/r2f_pytest.log; tail -c 400 gpurun_out/r2f_bench_ref.log
for f in r2f_bench r2f_bench_nooverlap; do python - <<PY
import json
for l in open('gpurun_out/$f.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$f', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), round(d['roofline']['step_tensor_frac'],3), d.get('cpu_baseline'))
PY
done
