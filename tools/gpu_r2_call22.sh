#!/bin/bash
# Round 2, call 22 (one B200): implicit-GEMM convolution weight gradient (halo views and per-tap boxes), training step.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider -k "implicit_gemm" -s > $OUT/pytest_dw_halo.log 2>&1; echo "pytest dw (halo) rc=$?"
tail -3 $OUT/pytest_dw_halo.log; grep -E "FAILED|Error|\[train-op\]" $OUT/pytest_dw_halo.log | head -20
RELPOSE_DW_HALO=0 timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider -k "implicit_gemm" -s > $OUT/pytest_dw_taps.log 2>&1; echo "pytest dw (per-tap boxes) rc=$?"
tail -3 $OUT/pytest_dw_taps.log; grep -E "FAILED|Error|\[train-op\]" $OUT/pytest_dw_taps.log | head -20
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider -s > $OUT/pytest_train.log 2>&1; echo "pytest train rc=$?"
tail -3 $OUT/pytest_train.log; grep -E "FAILED|Error|assert" $OUT/pytest_train.log | head -20
timeout 300 python -m rel_pose_b200.train_synthetic --steps 20 --warmup_steps 5 > $OUT/train_graph_c22.json 2> $OUT/train_graph_c22.err; echo "train graph rc=$?"; tail -3 $OUT/train_graph_c22.err; head -c 1200 $OUT/train_graph_c22.json; echo
RELPOSE_DW_HALO=0 timeout 300 python -m rel_pose_b200.train_synthetic --steps 20 --warmup_steps 5 > $OUT/train_graph_c22_taps.json 2> $OUT/train_graph_c22_taps.err; echo "train graph (per-tap) rc=$?"; head -c 400 $OUT/train_graph_c22_taps.json; echo
