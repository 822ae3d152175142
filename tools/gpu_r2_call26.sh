#!/bin/bash
# Round 2, call 26 (one B200): two-pass max-pool gradient, vectorised column sums, 16-lane BatchNorm finals, 32-row LayerNorm gradient blocks.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider -s > $OUT/pytest_train.log 2>&1; echo "pytest train rc=$?"
tail -3 $OUT/pytest_train.log; grep -E "FAILED|Error|assert" $OUT/pytest_train.log | head -20
timeout 300 python -m rel_pose_b200.train_synthetic --steps 20 --warmup_steps 5 > $OUT/train_graph_c26.json 2> $OUT/train_graph_c26.err; echo "train graph rc=$?"; tail -3 $OUT/train_graph_c26.err; head -c 700 $OUT/train_graph_c26.json; echo
