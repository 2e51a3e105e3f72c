#!/bin/bash
mkdir -p gpurun_out
echo "== staged"; python tools/ew_precision.py
echo "== rows"; MOPOE_EW_STAGED=0 python tools/ew_precision.py
echo "== old"; MOPOE_EW_STAGED=0 MOPOE_EW_ROWS=0 python tools/ew_precision.py
T="tests/test_gpu_parity.py::test_fp32_gradients_of_smooth_loss_match_oracle"
echo "== default"; timeout 600 python -m pytest $T -q 2>&1 | grep -E "^FAILED|passed|failed|assert [0-9]+ <="
echo "== nofuse";  MOPOE_FUSE_NEXT_BN_STATS=0 timeout 600 python -m pytest $T -q 2>&1 | grep -E "^FAILED|passed|failed|assert [0-9]+ <="
echo "== k1";  MOPOE_LIB_PATH=$PWD/tools/variants/lib_k1.so MOPOE_FUSE_NEXT_BN_STATS=0 timeout 600 python -m pytest $T -q 2>&1 | grep -E "^FAILED|passed|failed|assert [0-9]+ <="
echo "== rows";  MOPOE_EW_STAGED=0 timeout 600 python -m pytest $T -q 2>&1 | grep -E "^FAILED|passed|failed|assert [0-9]+ <="
echo "== old";  MOPOE_EW_STAGED=0 MOPOE_EW_ROWS=0 timeout 600 python -m pytest $T -q 2>&1 | grep -E "^FAILED|passed|failed|assert [0-9]+ <="
