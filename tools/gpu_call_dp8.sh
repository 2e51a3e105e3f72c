#!/bin/bash
# N-GPU call: weak scaling with / without the bucketed overlap, BASELINE configs 3 (strong), 4 and 5
mkdir -p gpurun_out
N=${1:-8}
shift
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
P=29600
run() {  # name, extra args...
  name=$1; shift
  P=$((P+1))
  timeout 500 $TR --master-port $P bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r2_n${N}_$name.log 2> gpurun_out/r2_n${N}_$name.err
  echo "$name exit $?"
}
for what in "$@"; do
  case $what in
    weak) run weak ;;
    weak_nobuckets) MOPOE_DP_BUCKETS=0 run weak_nobuckets ;;
    cfg3) run cfg3 --scaling strong --global-batch 2048 ;;
    cfg4) run cfg4 --config 4 ;;
    cfg5-joint) run cfg5-joint --config 5-joint ;;
    cfg5-moe) run cfg5-moe --config 5-moe ;;
    cfg5-poe) run cfg5-poe --config 5-poe ;;
    dptest) timeout 600 python -m pytest tests/test_gpu_dp.py -q -x > gpurun_out/r2_dp_pytest_n$N.log 2>&1; echo "dp pytest exit $?"; tail -n 3 gpurun_out/r2_dp_pytest_n$N.log ;;
  esac
done
for f in gpurun_out/r2_n${N}_*.log; do python - <<PY
import json
for l in open('$f'):
    if l.startswith('{'):
        d=json.loads(l); dc=d.get('dp_check') or {}
        print('$f', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'dp_check', dc.get('max_abs_err'), dc.get('params_equal_across_ranks'), dc.get('multicast'))
PY
done
