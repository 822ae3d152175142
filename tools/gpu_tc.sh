#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -rA -p no:cacheprovider -x > $OUT/pytest_tc.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_tc.log
grep -E "parity|tc-diag|FAILED|passed|failed|Error" $OUT/pytest_tc.log | head -60
