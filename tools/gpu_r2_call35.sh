#!/bin/bash
# Round 2, calls 35 / 36 (N = 2 / 4 B200s, `gpurun --gpus N -- bash tools/gpu_r2_call35.sh N`): the final build's bench under
# torchrun -- headline (weak scaling), config 4 (4096 pairs in total, strong scaling), config 5 (training step with the gradient
# exchange inside the captured step).
set -u
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 \
    bench.py --gpus $N --steps 20 --warmup 5 --legs main,config4,config5 > $OUT/bench_${N}gpu_c35.json 2> $OUT/bench_${N}gpu_c35.err; echo "bench ${N}gpu rc=$?"
tail -2 $OUT/bench_${N}gpu_c35.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("$OUT/bench_${N}gpu_c35.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"] if d.get("e2e") else None,"n",d["n_gpus"],"ms",d["ms_per_step"], "clocks", d["clocks"])
for k in ("config4","config5"):
    print(k, json.dumps(d.get(k))[:900])
PY
