#!/bin/bash
# GPU-box visit while iterating on the fused MLP kernel: its parity test, forward parity, bench with / without it.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -rA -p no:cacheprovider -k "${1:-mlp_fused}" > $OUT/pytest_mlp.log 2>&1; echo "mlp rc=$?" | tee -a $OUT/pytest_mlp.log
grep -E "parity\] mlp|tc-diag|FAILED|passed|failed|Error|error" $OUT/pytest_mlp.log | head -40
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q -rA -p no:cacheprovider > $OUT/pytest_fwd.log 2>&1; echo "fwd rc=$?"
grep -E "parity\].*(precision|rot_err)|FAILED|passed|failed|Error" $OUT/pytest_fwd.log | head -30
for cfg in ${BENCH_CFGS:-bf16x3:1 bf16:1}; do
  set -- ${cfg%%:*} ${cfg##*:}
  RELPOSE_FUSED_MLP=$2 timeout 600 python bench.py --steps 10 --warmup 3 --precision $1 --no-cpu-baseline > $OUT/bench_$1_f$2.json 2> $OUT/bench_$1_f$2.err; echo "bench $1 fused=$2 rc=$?"; tail -3 $OUT/bench_$1_f$2.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_$1_f$2.json"))
print("$1 fused=$2 value",round(d['value'],1),'e2e',round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],3))
for k,v in list(d['stages'].items())[:14]: print(f"  {k:28s} {v['calls']:3d} {v['ms']:8.3f} ms {100*v['share']:5.1f}% {v['tflops']:7.2f} TF {v['gbs']:7.1f} GB/s")
PY
done
