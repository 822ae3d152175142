#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -rA -p no:cacheprovider -k "self_attention" > $OUT/pytest_attn.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_attn.log
grep -E "parity\]|tc-diag|FAILED|passed|failed|Error|error" $OUT/pytest_attn.log | head -60
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q -rA -p no:cacheprovider -k "tensor_core" > $OUT/pytest_fwd_tc.log 2>&1; echo "rc=$?"
grep -E "parity\].*precision|FAILED|passed|failed|Error" $OUT/pytest_fwd_tc.log | head -30
for prec in bf16x3 bf16; do
  timeout 600 python bench.py --steps 10 --warmup 3 --precision $prec --no-cpu-baseline > $OUT/bench_$prec.json 2> $OUT/bench_$prec.err; echo "bench $prec rc=$?"; tail -3 $OUT/bench_$prec.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_$prec.json"))
print("$prec value",round(d['value'],1),'e2e',round(d['e2e']['value'],1))
for k,v in list(d['stages'].items())[:14]: print(f"  {k:28s} {v['calls']:3d} {v['ms']:8.3f} ms {100*v['share']:5.1f}% {v['tflops']:7.2f} TF {v['gbs']:7.1f} GB/s")
PY
done
