#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q -rA -p no:cacheprovider > $OUT/pytest_fwd.log 2>&1; echo "fwd rc=$?"
grep -E "FAILED|passed|failed|Error|error:|assert " $OUT/pytest_fwd.log | cut -c1-220 | head -20
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_bf16x3.json 2> $OUT/bench_bf16x3.err; echo "bench rc=$?"; tail -3 $OUT/bench_bf16x3.err
python - <<PY
import json
d=json.load(open("$OUT/bench_bf16x3.json"))
print("value",round(d['value'],1),'ms/step',round(d['ms_per_step'],3)); print('e2e',d['e2e']); print('e2e_u8',d['e2e_u8']); print('roofline',d['roofline'])
PY
