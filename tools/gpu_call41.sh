#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r4b_suite.log 2>&1; echo "suite exit $?" >> gpurun_out/r4b_suite.log
grep -E "^FAILED|passed|failed|exit|^E " gpurun_out/r4b_suite.log | head -40
