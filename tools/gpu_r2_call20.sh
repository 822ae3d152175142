#!/bin/bash
# Round 2, call 20 (one B200): flash-style attention forward (log-sum-exp) + backward on tcgen05 in the training step.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x -p no:cacheprovider -k "attention" -s > $OUT/pytest_attn_bwd.log 2>&1; echo "pytest attention rc=$?"
tail -3 $OUT/pytest_attn_bwd.log; grep -E "FAILED|Error|assert|\[train-op\]" $OUT/pytest_attn_bwd.log | head -30
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x -p no:cacheprovider -s > $OUT/pytest_train.log 2>&1; echo "pytest train rc=$?"
tail -3 $OUT/pytest_train.log; grep -E "FAILED|Error|assert" $OUT/pytest_train.log | head -20
timeout 300 python -m rel_pose_b200.train_synthetic --steps 20 --warmup_steps 5 > $OUT/train_graph_c20.json 2> $OUT/train_graph_c20.err; echo "train graph rc=$?"; tail -3 $OUT/train_graph_c20.err; head -c 1200 $OUT/train_graph_c20.json; echo
RELPOSE_TRAIN_FLASH=0 timeout 300 python -m rel_pose_b200.train_synthetic --steps 20 --warmup_steps 5 > $OUT/train_graph_c20_noflash.json 2> $OUT/train_graph_c20_noflash.err; echo "train graph (materialised) rc=$?"; head -c 700 $OUT/train_graph_c20_noflash.json; echo
