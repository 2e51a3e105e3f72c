#!/bin/bash
mkdir -p gpurun_out
MOPOE_GEMM_SPLITK=1 timeout 600 python -m pytest tests/test_gpu_gemm.py -q -k "split_k" 2>&1 | tail -3
MOPOE_GEMM_SPLITK=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q --maxfail=5 2>&1 | tail -3
# opt-in / fallback switches still work end to end
for sw in MOPOE_EW_STAGED=0 MOPOE_GEMM_PAIR=0 MOPOE_WGRAD_PAIR=0 MOPOE_GATE_RECOMPUTE=0 MOPOE_WGRAD_STREAMS=1 MOPOE_FUSE_NEXT_BN_STATS=0; do
  env $sw timeout 600 python -m pytest tests/test_gpu_parity.py -q --maxfail=3 -k "tri_joint or tri_moe or patext" 2>&1 | tail -1 | sed "s/^/$sw: /"
done
# a longer run: 300 steps
timeout 900 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/r3k_bench_300.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/r3k_bench_300.log'):
    if l.startswith('{'):
        d=json.loads(l); print('300 steps', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'], d['last_step'])
PY
