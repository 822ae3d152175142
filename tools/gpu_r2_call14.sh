#!/bin/bash
# Round 2, call 14 (EIGHT B200s): bench at N = 8 (headline weak scaling + config 4 strong scaling + config 5 legs),
# config 5 by itself (CUDA graphs + flat all-reduce, and the reference's DistributedDataParallel loop with the exchange timings).
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT/gpus_call14.txt; nvidia-smi topo -m > $OUT/topo_call14.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --gpus 8 --steps 10 --warmup 3 > $OUT/bench_8gpu.json 2> $OUT/bench_8gpu.err; echo "bench 8gpu rc=$?"
tail -2 $OUT/bench_8gpu.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29572 \
    -m rel_pose_b200.train_synthetic --steps 30 --warmup_steps 5 > $OUT/train_8gpu_graph.json 2> $OUT/train_8gpu_graph.err; echo "train 8gpu graph rc=$?"
tail -c 1300 $OUT/train_8gpu_graph.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29573 \
    -m rel_pose_b200.train_synthetic --steps 30 --warmup_steps 5 --graph 0 > $OUT/train_8gpu_ddp.json 2> $OUT/train_8gpu_ddp.err; echo "train 8gpu ddp rc=$?"
tail -c 1300 $OUT/train_8gpu_ddp.json; echo
python - <<PY
import json
d=json.load(open("$OUT/bench_8gpu.json"))
print("value",d["value"],"e2e",d["e2e"],"n",d["n_gpus"],"ms",d["ms_per_step"])
for k in ("config4","config5","clocks"):
    print(k, json.dumps(d.get(k))[:1100])
PY
