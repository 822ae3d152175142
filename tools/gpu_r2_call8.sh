#!/bin/bash
# Round 2, call 8 (one B200): geometry kernels (carried norms), new entry-point tests, line-level ncu captures of the hot kernels.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_tc.py -m gpu -q -x -p no:cacheprovider > $OUT/pytest_tc.log 2>&1; echo "pytest ops+tc rc=$?"
tail -3 $OUT/pytest_tc.log; grep -E "FAILED|Error|assert" $OUT/pytest_tc.log | head -20
timeout 600 python bench.py --legs main,geometry --no-cpu-baseline > $OUT/bench_c8.json 2> $OUT/bench_c8.err; echo "bench rc=$?"; tail -3 $OUT/bench_c8.err
BENCH="python bench.py --steps 1 --warmup 1 --legs main --no-cpu-baseline --no-e2e"
for k in mlp_fused_tc_kernel self_attention_tc_kernel ln_linear_tc_kernel conv3x3_halo_tc_kernel em_accum2_tc_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $OUT/r2c8_$k $BENCH > $OUT/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
python - <<PY
import json
d=json.load(open("$OUT/bench_c8.json"))
print("value",round(d['value'],1),'e2e',d['e2e'] and round(d['e2e']['value'],1),'clocks',d['clocks'])
for k,v in list(d['stages'].items())[:12]: print(f"  {k:32s} {v['calls']:3d} {v['ms']:8.3f} ms {100*v['share']:5.1f}% {v['tflops']:7.2f} TF")
print(json.dumps(d.get('geometry'))[:1500])
PY
