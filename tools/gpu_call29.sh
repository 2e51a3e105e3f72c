#!/bin/bash
# round-2 late state: smoke, every BASELINE config on one GPU, 8-bit wire formats, reference arm, in-graph kernel times, GEMM shapes
mkdir -p gpurun_out
timeout 400 python __graft_entry__.py smoke > gpurun_out/r3a_smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/r3a_smoke.log
for c in 4 5-joint 5-moe 5-poe; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3a_bench_cfg$c.log 2> gpurun_out/r3a_bench_cfg$c.err; echo "cfg $c exit $?"
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --text-wire uint8 --image-wire uint8 > gpurun_out/r3a_bench_wire_u8.log 2> gpurun_out/r3a_bench_wire_u8.err; echo "wire exit $?"
MOPOE_BENCH_SHAPES=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3a_bench.log 2> gpurun_out/r3a_bench_shapes.log; echo "bench (with cpu baseline) exit $?"
MOPOE_BRANCH_STREAMS=0 timeout 300 python bench.py --steps 3 --warmup 5 --no-cpu-baseline --profile-kernels > gpurun_out/r3a_in_graph_kernel_times.txt 2>&1
for f in r3a_bench r3a_bench_cfg4 r3a_bench_cfg5-joint r3a_bench_cfg5-moe r3a_bench_cfg5-poe r3a_bench_wire_u8; do python - <<PY
import json
for l in open('gpurun_out/$f.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$f', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'h2d', d['e2e']['h2d_bytes_per_step'], 'frac', round(d['roofline']['frac'],3), round(d['roofline']['step_tensor_frac'],3), d['last_step'], d.get('cpu_baseline'))
PY
done
sed -n 3,30p gpurun_out/r3a_in_graph_kernel_times.txt | cut -c1-100
