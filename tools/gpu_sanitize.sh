#!/bin/bash
# compute-sanitizer over the kernels added at the end of round 2 (small shapes): memcheck on em_project v2, the cluster regressor
# tail (distributed shared memory), the batched split-K reduce and the dead-row paths of attention / the Essential Matrix Module;
# racecheck on the two SIMT kernels that exchange data through shared memory.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "em_project_versions or regressor_tail" \
  > $OUT/sanitize_memcheck_ops.log 2>&1; echo "memcheck ops rc=$?"
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_tc.py -m gpu -q -x -p no:cacheprovider -k "(self_attention_tc or essential_tc or linear_tc_splitk) and not 26880 and not 24768" \
  > $OUT/sanitize_memcheck_tc.log 2>&1; echo "memcheck tc rc=$?"
timeout 120 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "em_project_versions or regressor_tail_cluster" \
  > $OUT/sanitize_racecheck_ops.log 2>&1; echo "racecheck ops rc=$?"
for f in $OUT/sanitize_memcheck_ops.log $OUT/sanitize_memcheck_tc.log $OUT/sanitize_racecheck_ops.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|out of bounds|misaligned|hazard" $f | head -8; done
