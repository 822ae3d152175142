#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_train.py -m gpu -q -rA -p no:cacheprovider ${1:+-k "$1"} > $OUT/pytest_train.log 2>&1; echo "train rc=$?"
grep -E "^\[train|FAILED|passed|failed|Error|error:|assert " $OUT/pytest_train.log | cut -c1-260 | head -80
