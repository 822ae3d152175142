#!/bin/bash
# --noess / CNN-only ablations: cross-attention kernels, pooling-head GEMM shapes, whole-path parity in the three precisions
set -u
OUT=gpurun_out; mkdir -p $OUT
K=${1:-"noess or cnn_only"}
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -p no:cacheprovider -k "cross_attention or 96-384 or 24768" > $OUT/pytest_noess_tc.log 2>&1; echo "tc rc=$?"; tail -3 $OUT/pytest_noess_tc.log
timeout 600 python -m pytest tests/test_gpu_forward.py -m gpu -q -s -p no:cacheprovider -k "$K" > $OUT/pytest_noess_fwd.log 2>&1; echo "fwd rc=$?"; grep "parity\] ablate" $OUT/pytest_noess_fwd.log; tail -3 $OUT/pytest_noess_fwd.log
grep -E "FAILED|Error" $OUT/pytest_noess_tc.log $OUT/pytest_noess_fwd.log | head -20
