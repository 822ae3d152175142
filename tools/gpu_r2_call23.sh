#!/bin/bash
# Round 2, call 23 (one B200): stride-2 implicit dW, vectorised BatchNorm passes with the ReLU mask fused, stem im2col written
# as transposed planes, split-K regressor forward; launch list of the new step.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider -s > $OUT/pytest_train.log 2>&1; echo "pytest train rc=$?"
tail -3 $OUT/pytest_train.log; grep -E "FAILED|Error|assert" $OUT/pytest_train.log | head -20; grep -E "conv dW.*/s2|golden|\[train\] " $OUT/pytest_train.log | cut -c1-220 | head
timeout 300 python -m rel_pose_b200.train_synthetic --steps 20 --warmup_steps 5 > $OUT/train_graph_c23.json 2> $OUT/train_graph_c23.err; echo "train graph rc=$?"; tail -3 $OUT/train_graph_c23.err; head -c 900 $OUT/train_graph_c23.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 1000 --csv --log-file $OUT/train_launches_c23.csv \
    python -m rel_pose_b200.train_synthetic --steps 2 --warmup_steps 2 --batch 6 --pool 2 --graph 0 > $OUT/ncu_train.log 2>&1; echo "train list rc=$?"
