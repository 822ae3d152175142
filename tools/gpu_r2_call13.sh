#!/bin/bash
# Round 2, call 13 (TWO B200s): config 5 as two CUDA graphs with one flat-gradient NCCL all-reduce between them.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x -p no:cacheprovider -k "cuda_graph" -s > $OUT/pytest_train.log 2>&1; echo "pytest train rc=$?"
tail -3 $OUT/pytest_train.log; grep -E "FAILED|Error|assert|\[train\]" $OUT/pytest_train.log | head -10
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
    -m rel_pose_b200.train_synthetic --steps 20 --warmup_steps 5 > $OUT/train_2gpu_graph2.json 2> $OUT/train_2gpu_graph2.err; echo "train 2gpu graph rc=$?"
tail -c 1500 $OUT/train_2gpu_graph2.json; tail -3 $OUT/train_2gpu_graph2.err | cut -c1-300
