#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_train.py -m gpu -q -rA -p no:cacheprovider > $OUT/pytest_train.log 2>&1; echo "train rc=$?"
grep -E "^\[train|FAILED|passed|failed|Error|error:|assert " $OUT/pytest_train.log | cut -c1-220 | head -30
timeout 600 python -m rel_pose_b200.train_synthetic --steps 10 --warmup_steps 3 --batch 6 > $OUT/train_fused.json 2> $OUT/train_fused.err; echo "train rc=$?"; tail -3 $OUT/train_fused.err; cat $OUT/train_fused.json
