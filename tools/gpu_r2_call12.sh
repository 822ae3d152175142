#!/bin/bash
# Round 2, call 12 (TWO B200s): config 5 under DDP with the step captured as one CUDA graph (falls back to eager if the
# collective cannot be captured), and the eager DDP loop with the exchange measurements.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
    -m rel_pose_b200.train_synthetic --steps 20 --warmup_steps 12 > $OUT/train_2gpu_graph.json 2> $OUT/train_2gpu_graph.err; echo "train 2gpu graph rc=$?"
tail -c 1500 $OUT/train_2gpu_graph.json; tail -5 $OUT/train_2gpu_graph.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 \
    -m rel_pose_b200.train_synthetic --steps 20 --warmup_steps 5 --graph 0 > $OUT/train_2gpu_eager.json 2> $OUT/train_2gpu_eager.err; echo "train 2gpu eager rc=$?"
tail -c 1200 $OUT/train_2gpu_eager.json
