#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_elementwise.py -q --maxfail=50 > gpurun_out/r2k_ew_pytest.log 2>&1; echo "ew pytest exit $?" >> gpurun_out/r2k_ew_pytest.log
grep -E "^FAILED|passed|failed|exit" gpurun_out/r2k_ew_pytest.log | head -20
echo "== staged"; timeout 300 python tools/prof_ew_shapes.py 2>&1 | tee gpurun_out/r2k_shapes_staged.txt
echo "== rows"; MOPOE_EW_STAGED=0 timeout 300 python tools/prof_ew_shapes.py 2>&1 | tee gpurun_out/r2k_shapes_rows.txt
echo "== old"; MOPOE_EW_STAGED=0 MOPOE_EW_ROWS=0 timeout 300 python tools/prof_ew_shapes.py 2>&1 | tee gpurun_out/r2k_shapes_old.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2k_bench.log 2>&1; echo "bench exit $?"
python - <<PY
import json
for l in open('gpurun_out/r2k_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print('r2k', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), round(d['roofline']['step_tensor_frac'],3), {k:(round(v['ms'],2),round(v['frac'],2)) for k,v in d['roofline_hbm']['classes'].items()})
PY
