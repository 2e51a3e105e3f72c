#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r4j_bench.log 2> gpurun_out/r4j_bench.err; echo "bench exit $?"
python - <<PY
import json
for l in open('gpurun_out/r4j_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['roofline']['fused_epilogues'], 'frac', round(d['roofline']['frac'],3), d['cpu_baseline']['value'])
PY
