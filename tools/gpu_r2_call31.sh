#!/bin/bash
# Round 2, call 31 (TWO B200s): config 5 with the gradient exchange inside the captured step (two buckets, the large one under the
# CNN's backward) against the single all-reduce between the graphs: same losses, step time, phases.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 \
    -m rel_pose_b200.train_synthetic --steps 30 --warmup_steps 5 > $OUT/train_2gpu_overlap_c31.json 2> $OUT/train_2gpu_overlap_c31.err; echo "train 2gpu overlap rc=$?"
tail -3 $OUT/train_2gpu_overlap_c31.err | cut -c1-400; tail -c 1800 $OUT/train_2gpu_overlap_c31.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29582 \
    -m rel_pose_b200.train_synthetic --steps 30 --warmup_steps 5 --overlap_exchange 0 > $OUT/train_2gpu_single_c31.json 2> $OUT/train_2gpu_single_c31.err; echo "train 2gpu single rc=$?"
tail -c 1500 $OUT/train_2gpu_single_c31.json; echo
timeout 600 python -m pytest tests/test_gpu_ddp.py -m gpu -q -p no:cacheprovider > $OUT/pytest_ddp_c31.log 2>&1; echo "pytest ddp rc=$?"; tail -2 $OUT/pytest_ddp_c31.log
