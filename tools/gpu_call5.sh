#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r2e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2e_pytest.log
for c in 4 5-joint 5-moe 5-poe; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_cfg$c.log 2> gpurun_out/r2e_bench_cfg$c.err; echo "cfg $c exit $?"
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --text-wire uint8 --image-wire uint8 > gpurun_out/r2e_bench_wire_u8.log 2> gpurun_out/r2e_bench_wire_u8.err; echo "wire exit $?"
timeout 400 python __graft_entry__.py smoke > gpurun_out/r2e_smoke.log 2>&1; echo "smoke exit $?"
tail -n 4 gpurun_out/r2e_pytest.log gpurun_out/r2e_smoke.log
for f in r2e_bench_cfg4 r2e_bench_cfg5-joint r2e_bench_cfg5-moe r2e_bench_cfg5-poe r2e_bench_wire_u8; do python - <<PY
import json
for l in open('gpurun_out/$f.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$f', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'h2d', d['e2e']['h2d_bytes_per_step'], 'step_tensor_frac', round(d['roofline']['step_tensor_frac'],3), d['last_step'])
PY
done
