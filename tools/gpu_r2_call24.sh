#!/bin/bash
# Round 2, call 24 (one B200): flash-style Essential Matrix Module backward (em_bwd_tc.cu), stem convolution on the window planes.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider -k "essential or stem_conv" -s > $OUT/pytest_em_bwd.log 2>&1; echo "pytest em+stem rc=$?"
tail -3 $OUT/pytest_em_bwd.log; grep -E "FAILED|Error|\[train-op\]" $OUT/pytest_em_bwd.log | cut -c1-200 | head -40
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider -s > $OUT/pytest_train.log 2>&1; echo "pytest train rc=$?"
tail -3 $OUT/pytest_train.log; grep -E "FAILED|Error|assert" $OUT/pytest_train.log | head -20; grep -E "\[train\] train_b" $OUT/pytest_train.log | cut -c1-200 | head
timeout 300 python -m rel_pose_b200.train_synthetic --steps 20 --warmup_steps 5 > $OUT/train_graph_c24.json 2> $OUT/train_graph_c24.err; echo "train graph rc=$?"; tail -3 $OUT/train_graph_c24.err; head -c 900 $OUT/train_graph_c24.json; echo
