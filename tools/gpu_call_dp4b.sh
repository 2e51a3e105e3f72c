#!/bin/bash
# 4-GPU call (late round 2): DP parity test, weak-scaling bench with dp_check (NVSwitch multicast exchange), config 3 @ 4 GPUs
mkdir -p gpurun_out
N=${1:-4}
timeout 600 python -m pytest tests/test_gpu_dp.py -q -x > gpurun_out/r3c_dp_pytest_n$N.log 2>&1; echo "dp pytest exit $?" >> gpurun_out/r3c_dp_pytest_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3c_bench_weak_n$N.log 2> gpurun_out/r3c_bench_weak_n$N.err; echo "exit $?" >> gpurun_out/r3c_bench_weak_n$N.err
timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --scaling strong --global-batch 2048 > gpurun_out/r3c_bench_cfg3_n$N.log 2> gpurun_out/r3c_bench_cfg3_n$N.err; echo "exit $?" >> gpurun_out/r3c_bench_cfg3_n$N.err
tail -n 3 gpurun_out/r3c_dp_pytest_n$N.log; tail -n 1 gpurun_out/r3c_bench_weak_n$N.err gpurun_out/r3c_bench_cfg3_n$N.err
for f in r3c_bench_weak_n$N r3c_bench_cfg3_n$N; do python - <<PY
import json
for l in open('gpurun_out/$f.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$f', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d.get('dp_check'))
PY
done
