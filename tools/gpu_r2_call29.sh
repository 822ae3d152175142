#!/bin/bash
# Round 2, call 29 (one B200): final single-GPU evidence of the round: full GPU test suite, the default bench line, the ncu
# launch list of one bench step, one `ncu --set full` capture of the fused MLP as it is now, the training step.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_gpu_c29.log 2>&1; echo "pytest gpu rc=$?"; tail -3 $OUT/pytest_gpu_c29.log; grep -E "FAILED|Error" $OUT/pytest_gpu_c29.log | head
timeout 900 python bench.py > $OUT/bench_c29_default.json 2> $OUT/bench_c29_default.err; echo "bench rc=$?"; tail -2 $OUT/bench_c29_default.err | cut -c1-300
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_c29_reference_arm.json 2> $OUT/bench_c29_reference_arm.err; echo "bench reference rc=$?"; head -c 400 $OUT/bench_c29_reference_arm.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c29.csv python bench.py --steps 2 --warmup 1 --legs main --no-cpu-baseline --no-e2e > $OUT/bench_under_ncu.log 2>&1; echo "ncu launch list rc=$?"
BENCH="python bench.py --steps 1 --warmup 1 --legs main --no-cpu-baseline --no-e2e"
for k in mlp_fused_tc_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $OUT/r2c29_$k $BENCH > $OUT/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
timeout 300 python -m rel_pose_b200.train_synthetic --steps 20 --warmup_steps 5 > $OUT/train_graph_c29.json 2> $OUT/train_graph_c29.err; echo "train graph rc=$?"; head -c 700 $OUT/train_graph_c29.json; echo
python - <<PY
import json
d=json.loads(open("$OUT/bench_c29_default.json").read().strip().splitlines()[-1])
print("value",round(d['value'],1),'e2e',d['e2e'] and round(d['e2e']['value'],1),'clocks',d['clocks'])
for k,v in list(d['stages'].items())[:14]: print(f"  {k:32s} {v['calls']:3d} {v['ms']/v['calls']*1000:8.1f} us {100*v['share']:5.1f}% {v['tflops']:7.2f} TF")
for k in ('roofline','parity','gpu_eager_baseline','config4','config5','geometry','cpu_baseline'):
    print(k, json.dumps(d.get(k))[:600])
PY
