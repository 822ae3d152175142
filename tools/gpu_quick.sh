#!/bin/bash
# quick regression after a change in tc_common: every tensor-core op test, forward parity, bench
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -p no:cacheprovider > $OUT/pytest_tc.log 2>&1; echo "tc rc=$?"; tail -2 $OUT/pytest_tc.log
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q -p no:cacheprovider > $OUT/pytest_fwd.log 2>&1; echo "fwd rc=$?"; tail -2 $OUT/pytest_fwd.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_bf16x3.json 2> $OUT/bench_bf16x3.err; echo "bench rc=$?"; tail -2 $OUT/bench_bf16x3.err
python - <<PY
import json
d=json.load(open("$OUT/bench_bf16x3.json"))
print("value",round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],3))
for k,v in list(d['stages'].items())[:14]: print(f"  {k:32s} {v['calls']:3d} {v['ms']:8.3f} ms {100*v['share']:5.1f}% {v['tflops']:7.2f} TF")
PY
