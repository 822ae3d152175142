#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -rA -p no:cacheprovider -k "linear_tc" > $OUT/pytest_lin.log 2>&1; echo "lin rc=$?"
grep -E "parity\] linear_tc_splitk|FAILED|passed|failed|Error|error" $OUT/pytest_lin.log | cut -c1-200 | head -20
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q -rA -p no:cacheprovider > $OUT/pytest_fwd.log 2>&1; echo "fwd rc=$?"; tail -2 $OUT/pytest_fwd.log
grep -E "parity\].*precision=bf16x3" $OUT/pytest_fwd.log | cut -c1-160
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_chk.json 2> $OUT/bench_chk.err; echo "bench rc=$?"; tail -2 $OUT/bench_chk.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_chk.json").read())
print("value",round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],3),'clocks',d['clocks'])
for k,v in d['stages'].items():
    if 'linear' in k: print(f"  {k:32s} {v['calls']:3d} {v['ms']:8.3f} ms {100*v['share']:5.1f}% {v['tflops']:7.2f} TF {v['gbs']:7.1f} GB/s")
PY
wc -l $OUT/bench_chk.json
