#!/bin/bash
# 8-GPU call (end of round 2): weak-scaling bench with dp_check (NVSwitch multicast exchange), config 3 @ 8 GPUs = weak line
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3h_bench_weak_n$N.log 2> gpurun_out/r3h_bench_weak_n$N.err; echo "exit $?" >> gpurun_out/r3h_bench_weak_n$N.err
tail -n 1 gpurun_out/r3h_bench_weak_n$N.err
python - <<PY
import json
for l in open('gpurun_out/r3h_bench_weak_n8.log'):
    if l.startswith('{'):
        d=json.loads(l); print('weak n8', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d.get('dp_check'))
PY
