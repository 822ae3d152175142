#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -rA -p no:cacheprovider -k "conv2d_tc" > $OUT/pytest_conv.log 2>&1; echo "conv rc=$?"
grep -E "parity\] conv|tc-diag|FAILED|passed|failed|Error|error" $OUT/pytest_conv.log | cut -c1-200 | head -40
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q -p no:cacheprovider > $OUT/pytest_fwd.log 2>&1; echo "fwd rc=$?"; tail -2 $OUT/pytest_fwd.log
for h in 1; do
RELPOSE_CONV_HALO=$h timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_halo$h.json 2> $OUT/bench_halo$h.err; echo "bench halo=$h rc=$?"; tail -2 $OUT/bench_halo$h.err
python - <<PY
import json
d=json.load(open("$OUT/bench_halo$h.json"))
print("halo=$h value",round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],3))
for k,v in d['stages'].items():
    if 'conv' in k: print(f"  {k:32s} {v['calls']:3d} {v['ms']:8.3f} ms {100*v['share']:5.1f}% {v['tflops']:7.2f} TF")
PY
done
