#!/bin/bash
# final round-2 state: full GPU suite, smoke, bench (with the CPU reference arm), every config, in-graph kernel times, GEMM shapes,
# launch list, ncu --set full of the staged passes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r3g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3g_pytest.log
grep -E "^FAILED|passed|failed|exit" gpurun_out/r3g_pytest.log | head
timeout 400 python __graft_entry__.py smoke > gpurun_out/r3g_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/r3g_smoke.log | cut -c1-250
MOPOE_BENCH_SHAPES=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3g_bench.log 2> gpurun_out/r3g_bench_shapes.log; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3g_bench_ref.log 2>&1; echo "ref exit $?"
for c in 4 5-joint 5-moe 5-poe; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3g_bench_cfg$c.log 2> gpurun_out/r3g_bench_cfg$c.err; echo "cfg $c exit $?"
done
MOPOE_BRANCH_STREAMS=0 timeout 300 python bench.py --steps 3 --warmup 5 --no-cpu-baseline --profile-kernels > gpurun_out/r3g_in_graph_kernel_times.txt 2>&1
for f in r3g_bench r3g_bench_cfg4 r3g_bench_cfg5-joint r3g_bench_cfg5-moe r3g_bench_cfg5-poe; do python - <<PY
import json
for l in open('gpurun_out/$f.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$f', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), round(d['roofline']['step_tensor_frac'],3), 'hbm', round(d['roofline_hbm']['all']['frac'],3), d.get('cpu_baseline',{}) and d['cpu_baseline'].get('value'))
PY
done
tail -n 2 gpurun_out/r3g_bench_ref.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r3g_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/r3g_ncu_launches.log 2>&1; echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"staged_" -c 16 -o gpurun_out/r3g_staged python tools/prof_ew.py > gpurun_out/r3g_ncu_staged.log 2>&1; echo "ncu exit $?"
