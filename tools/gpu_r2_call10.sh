#!/bin/bash
# Round 2, call 10 (one B200): CUDA-graph training step, MLP epilogue changes, geometry after the cheaper guard.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x -p no:cacheprovider -k "cuda_graph or fused_adam or eval_after" -s > $OUT/pytest_train.log 2>&1; echo "pytest train rc=$?"
tail -3 $OUT/pytest_train.log; grep -E "FAILED|Error|assert|\[train\]" $OUT/pytest_train.log | head -20
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "mlp or proj or planes or svd or essential_to_rt" > $OUT/pytest_tc.log 2>&1; echo "pytest tc rc=$?"
tail -3 $OUT/pytest_tc.log; grep -E "FAILED|Error|assert" $OUT/pytest_tc.log | head -20
timeout 300 python tools/bench_geom.py > $OUT/geom_c10.json 2> $OUT/geom_c10.err; echo "bench geom rc=$?"; head -c 700 $OUT/geom_c10.json; echo
timeout 300 python -m rel_pose_b200.train_synthetic --steps 20 --warmup_steps 5 > $OUT/train_graph.json 2> $OUT/train_graph.err; echo "train graph rc=$?"; tail -3 $OUT/train_graph.err; head -c 1200 $OUT/train_graph.json; echo
timeout 300 python -m rel_pose_b200.train_synthetic --steps 20 --warmup_steps 5 --graph 0 > $OUT/train_eager.json 2> $OUT/train_eager.err; echo "train eager rc=$?"; head -c 600 $OUT/train_eager.json; echo
timeout 600 python bench.py --legs main,parity --no-cpu-baseline > $OUT/bench_c10.json 2> $OUT/bench_c10.err; echo "bench rc=$?"; tail -3 $OUT/bench_c10.err
python - <<PY
import json
d=json.load(open("$OUT/bench_c10.json"))
print("value",round(d['value'],1),'e2e',d['e2e'] and round(d['e2e']['value'],1),'clocks',d['clocks'])
for k,v in list(d['stages'].items())[:12]: print(f"  {k:32s} {v['calls']:3d} {v['ms']:8.3f} ms {100*v['share']:5.1f}% {v['tflops']:7.2f} TF")
print(json.dumps(d.get('parity'))[:400])
PY
