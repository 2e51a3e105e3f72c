#!/bin/bash
# compute-sanitizer over the kernel-level GPU tests (one tool per gpurun call: B200_PROFILING.md)
TOOL=${1:-memcheck}
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool $TOOL --error-exitcode 9 --launch-timeout 0 \
    python -m pytest tests/test_gpu_gemm.py tests/test_gpu_kernels.py -q -x -p no:cacheprovider \
    -k "not graph and not pipelined and not 4194304 and not 262144-71 and not 2900" \
    > gpurun_out/r2_sanitizer_$TOOL.log 2>&1
echo "compute-sanitizer $TOOL exit $?" >> gpurun_out/r2_sanitizer_$TOOL.log
grep -c "Invalid\|Race\|hazard" gpurun_out/r2_sanitizer_$TOOL.log
tail -n 12 gpurun_out/r2_sanitizer_$TOOL.log
