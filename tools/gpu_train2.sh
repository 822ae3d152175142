#!/bin/bash
# fused optimizer parity, training driver phase breakdown (fused vs torch optimizer), geometry microbench
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -rA -p no:cacheprovider -k "fused_adam" > $OUT/pytest_opt.log 2>&1; echo "opt rc=$?"
grep -E "FAILED|passed|failed|Error|error:|assert " $OUT/pytest_opt.log | cut -c1-260 | head -20
for o in fused torch; do
  timeout 600 python -m rel_pose_b200.train_synthetic --steps 10 --warmup_steps 3 --batch 6 --optimizer $o > $OUT/train_$o.json 2> $OUT/train_$o.err; echo "train $o rc=$?"; tail -3 $OUT/train_$o.err; cat $OUT/train_$o.json
done
timeout 300 python tools/bench_geom.py > $OUT/geom.json 2> $OUT/geom.err; echo "geom rc=$?"; tail -3 $OUT/geom.err; cat $OUT/geom.json
