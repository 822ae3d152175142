#!/bin/bash
# Round 2, call 21 (one B200): full training tests with the flash attention path + launch list of the training step.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider -s > $OUT/pytest_train.log 2>&1; echo "pytest train rc=$?"
tail -3 $OUT/pytest_train.log; grep -E "FAILED|Error|assert" $OUT/pytest_train.log | head -20
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 1200 --csv --log-file $OUT/train_launches_c21.csv \
    python -m rel_pose_b200.train_synthetic --steps 2 --warmup_steps 2 --batch 6 --pool 2 --graph 0 > $OUT/ncu_train.log 2>&1; echo "train list rc=$?"
wc -l $OUT/train_launches_c21.csv
