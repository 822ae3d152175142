#!/bin/bash
# One GPU-box visit while iterating on kernels: tensor-core op tests, forward parity, bench (both precisions).
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -rA -p no:cacheprovider ${1:+-k "$1"} > $OUT/pytest_tc.log 2>&1; echo "tc rc=$?" | tee -a $OUT/pytest_tc.log
grep -E "parity\] (stem|self_att|em_)|tc-diag|FAILED|passed|failed|Error|error" $OUT/pytest_tc.log | head -40
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q -rA -p no:cacheprovider > $OUT/pytest_fwd.log 2>&1; echo "fwd rc=$?"
grep -E "parity\].*(precision|rot_err)|FAILED|passed|failed|Error" $OUT/pytest_fwd.log | head -30
for prec in bf16x3 bf16; do
  timeout 600 python bench.py --steps 10 --warmup 3 --precision $prec --no-cpu-baseline > $OUT/bench_$prec.json 2> $OUT/bench_$prec.err; echo "bench $prec rc=$?"; tail -3 $OUT/bench_$prec.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_$prec.json"))
print("$prec value",round(d['value'],1),'e2e',round(d['e2e']['value'],1))
for k,v in list(d['stages'].items())[:16]: print(f"  {k:28s} {v['calls']:3d} {v['ms']:8.3f} ms {100*v['share']:5.1f}% {v['tflops']:7.2f} TF {v['gbs']:7.1f} GB/s")
PY
done
