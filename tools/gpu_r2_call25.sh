#!/bin/bash
# Round 2, call 25 (one B200): training tests per engine, config-5 leg with the reference's own eager training step beside it,
# launch list of the step.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider -s > $OUT/pytest_train.log 2>&1; echo "pytest train rc=$?"
tail -3 $OUT/pytest_train.log; grep -E "FAILED|Error|assert" $OUT/pytest_train.log | head -20; grep -E "\[train\] train_b" $OUT/pytest_train.log | cut -c1-170 | head -12
timeout 900 python bench.py --legs config5 --no-cpu-baseline > $OUT/bench_c25_config5.json 2> $OUT/bench_c25_config5.err; echo "bench config5 rc=$?"; tail -3 $OUT/bench_c25_config5.err
python - <<PY
import json
d=json.load(open("$OUT/bench_c25_config5.json"))
c=d.get('config5',{})
print({k:c.get(k) for k in ('ms_per_step','pairs_per_s','phase_ms','cuda_graph')})
print(json.dumps(c.get('gpu_eager_baseline'))[:900])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 900 --csv --log-file $OUT/train_launches_c25.csv \
    python -m rel_pose_b200.train_synthetic --steps 2 --warmup_steps 2 --batch 6 --pool 2 --graph 0 > $OUT/ncu_train.log 2>&1; echo "train list rc=$?"
