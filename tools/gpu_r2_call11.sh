#!/bin/bash
# Round 2, call 11 (one B200): shared-memory address space kept through the alignment (LDS/STS instead of generic LD/ST),
# optimizer scalar ring; full GPU test suite + bench.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"
tail -4 $OUT/pytest_gpu.log; grep -E "FAILED|Error|assert" $OUT/pytest_gpu.log | head -20
timeout 600 python bench.py --legs main,parity --no-cpu-baseline > $OUT/bench_c11.json 2> $OUT/bench_c11.err; echo "bench rc=$?"; tail -3 $OUT/bench_c11.err
python - <<PY
import json
d=json.load(open("$OUT/bench_c11.json"))
print("value",round(d['value'],1),'e2e',d['e2e'] and round(d['e2e']['value'],1),'clocks',d['clocks'])
for k,v in list(d['stages'].items())[:16]: print(f"  {k:32s} {v['calls']:3d} {v['ms']:8.3f} ms {100*v['share']:5.1f}% {v['tflops']:7.2f} TF")
print(json.dumps(d.get('parity'))[:300])
PY
