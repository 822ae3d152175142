#!/bin/bash
# Round 2, call 27 (one B200): nn.Linear weight gradients on the implicit-GEMM kernel (rp_linear_dw_tc), gradients assigned + one multi-tensor pack.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider -s > $OUT/pytest_train.log 2>&1; echo "pytest train rc=$?"
tail -3 $OUT/pytest_train.log; grep -E "FAILED|Error|assert|linear dW" $OUT/pytest_train.log | head -30
timeout 300 python -m rel_pose_b200.train_synthetic --steps 20 --warmup_steps 5 > $OUT/train_graph_c27.json 2> $OUT/train_graph_c27.err; echo "train graph rc=$?"; tail -3 $OUT/train_graph_c27.err; head -c 900 $OUT/train_graph_c27.json; echo
RELPOSE_TRAIN_LIN_DW_TC=0 timeout 300 python -m rel_pose_b200.train_synthetic --steps 20 --warmup_steps 5 > $OUT/train_graph_c27_splitk_ab.json 2> $OUT/train_graph_c27_splitk_ab.err; echo "train graph (split-K dW) rc=$?"; head -c 900 $OUT/train_graph_c27_splitk_ab.json; echo
