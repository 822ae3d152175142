#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider > $OUT/pytest_train.log 2>&1; echo "train rc=$?"; tail -2 $OUT/pytest_train.log
grep -E "^\[train\]" $OUT/pytest_train.log | cut -c1-200
timeout 600 python -m rel_pose_b200.train_synthetic --steps 20 --warmup_steps 5 --batch 6 > $OUT/train_fused.json 2> $OUT/train_fused.err; echo "train rc=$?"; tail -2 $OUT/train_fused.err; cut -c1-700 $OUT/train_fused.json
