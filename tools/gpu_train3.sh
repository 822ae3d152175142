#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -rA -p no:cacheprovider -k "fused_adam" > $OUT/pytest_opt.log 2>&1; echo "opt rc=$?"
grep -E "FAILED|passed|failed|Error|error:|assert " $OUT/pytest_opt.log | cut -c1-260 | head -20
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/train_launches.csv \
   python -m rel_pose_b200.train_synthetic --steps 1 --warmup_steps 1 --batch 6 --pool 1 > $OUT/ncu_train.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/ncu_train.log
