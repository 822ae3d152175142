#!/bin/bash
# Round 2, call 16 (one B200): two-rows-per-warp LayerNorm stage, SVD without V accumulation.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "svd or essential or se3" -s > $OUT/pytest_geom.log 2>&1; echo "pytest geom rc=$?"
tail -2 $OUT/pytest_geom.log; grep -E "FAILED|Error|assert|parity" $OUT/pytest_geom.log | head -12
timeout 300 python tools/bench_geom.py > $OUT/geom_c16.json 2> $OUT/geom_c16.err; echo "bench geom rc=$?"; head -c 600 $OUT/geom_c16.json; echo
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_forward.py -m gpu -q -x -p no:cacheprovider -s > $OUT/pytest_tcfwd.log 2>&1; echo "pytest tc+forward rc=$?"
tail -2 $OUT/pytest_tcfwd.log; grep -E "FAILED|Error|assert" $OUT/pytest_tcfwd.log | head -12
grep -E "\[parity\].*precision=bf16x3" $OUT/pytest_tcfwd.log > $OUT/parity_bf16x3_goldens.log; cat $OUT/parity_bf16x3_goldens.log | cut -c1-160
timeout 600 python bench.py --legs main,parity --no-cpu-baseline > $OUT/bench_c16.json 2> $OUT/bench_c16.err; echo "bench rc=$?"; tail -3 $OUT/bench_c16.err
python - <<PY
import json
d=json.load(open("$OUT/bench_c16.json"))
print("value",round(d['value'],1),'e2e',d['e2e'] and round(d['e2e']['value'],1),'clocks',d['clocks'])
for k,v in list(d['stages'].items())[:10]: print(f"  {k:32s} {v['calls']:3d} {v['ms']/v['calls']*1000:8.1f} us {100*v['share']:5.1f}% {v['tflops']:7.2f} TF")
print(json.dumps(d.get('parity'))[:300])
PY
