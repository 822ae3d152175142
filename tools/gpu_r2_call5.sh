#!/bin/bash
# Round 2, call 5 (TWO B200s): gradient exchange test (DDP mean of per-rank gradients), config 5 on 2 ranks, bench at N = 2.
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT/gpus_call5.txt
timeout 900 python -m pytest tests/test_gpu_ddp.py -m gpu -q -x -s -p no:cacheprovider > $OUT/pytest_ddp.log 2>&1; echo "pytest ddp rc=$?"
tail -8 $OUT/pytest_ddp.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    -m rel_pose_b200.train_synthetic --steps 20 --warmup_steps 5 > $OUT/train_2gpu.json 2> $OUT/train_2gpu.err; echo "train 2gpu rc=$?"
tail -c 1500 $OUT/train_2gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err; echo "bench 2gpu rc=$?"
tail -3 $OUT/bench_2gpu.err
python - <<PY
import json
d=json.load(open("$OUT/bench_2gpu.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"n",d["n_gpus"])
for k in ("config4","config5","clocks"):
    print(k, json.dumps(d.get(k))[:900])
PY
