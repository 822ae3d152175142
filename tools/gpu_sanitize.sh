#!/bin/bash
# compute-sanitizer memcheck over the session-3 kernels (small shapes) + stdout contract of bench.py
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_tc.py -m gpu -q -x -p no:cacheprovider \
  -k "(mlp_fused_tc or ln_linear_tc or self_attention_tc) and not 37965 and not 8960" \
  > $OUT/sanitize.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned" $OUT/sanitize.log | head -20
