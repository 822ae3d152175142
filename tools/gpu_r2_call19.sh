#!/bin/bash
# Round 2, call 19 (one B200): one elected tempty arrive per epilogue warp in gemm_tc (was 512 arrives per accumulator).
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_forward.py -m gpu -q -x -p no:cacheprovider -s > $OUT/pytest_tcfwd.log 2>&1; echo "pytest tc+forward rc=$?"
tail -2 $OUT/pytest_tcfwd.log; grep -E "FAILED|Error|assert" $OUT/pytest_tcfwd.log | head -12
timeout 600 python bench.py --legs main,parity --no-cpu-baseline > $OUT/bench_c19.json 2> $OUT/bench_c19.err; echo "bench rc=$?"; tail -3 $OUT/bench_c19.err
python - <<PY
import json
d=json.load(open("$OUT/bench_c19.json"))
print("value",round(d['value'],1),'e2e',d['e2e'] and round(d['e2e']['value'],1),'clocks',d['clocks'])
for k,v in list(d['stages'].items())[:14]: print(f"  {k:32s} {v['calls']:3d} {v['ms']/v['calls']*1000:8.1f} us {100*v['share']:5.1f}% {v['tflops']:7.2f} TF")
print(json.dumps(d.get('parity'))[:300])
PY
