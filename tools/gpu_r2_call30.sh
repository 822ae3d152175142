#!/bin/bash
# Round 2, call 30 (EIGHT B200s): the round's final multi-GPU numbers: bench at N = 8 (headline weak scaling, config 4 strong
# scaling, config 5 leg) and config 5 by itself (CUDA graphs + one flat-gradient NCCL all-reduce) with the current training step.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --gpus 8 --steps 10 --warmup 3 --legs main,config4,config5 > $OUT/bench_8gpu_c30.json 2> $OUT/bench_8gpu_c30.err; echo "bench 8gpu rc=$?"
tail -2 $OUT/bench_8gpu_c30.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29572 \
    -m rel_pose_b200.train_synthetic --steps 30 --warmup_steps 5 > $OUT/train_8gpu_c30.json 2> $OUT/train_8gpu_c30.err; echo "train 8gpu graph rc=$?"
tail -c 1500 $OUT/train_8gpu_c30.json; echo
python - <<PY
import json
d=json.loads(open("$OUT/bench_8gpu_c30.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"],"n",d["n_gpus"],"ms",d["ms_per_step"])
for k in ("config4","config5","clocks"):
    print(k, json.dumps(d.get(k))[:1100])
PY
