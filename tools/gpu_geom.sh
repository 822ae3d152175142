#!/bin/bash
# geometry kernels (config 3): parity tests + CUDA-graph microbench; staging-mode A/B on one box
set -u
OUT=gpurun_out; mkdir -p $OUT
for m in 0 1 2 0 1 2; do
  RELPOSE_GEOM_STAGE=$m timeout 300 python tools/bench_geom.py > $OUT/geom_mode$m.json 2> $OUT/geom.err; echo "mode $m rc=$?"
  python - <<PY
import json
d=json.load(open("$OUT/geom_mode$m.json"))["kernels"]
print("mode $m", {k: v["us"] for k,v in d.items()})
PY
done
