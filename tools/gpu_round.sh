#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list.  Everything under `timeout`.
# Usage (from the repo root on the box):  bash tools/gpu_round.sh [quick]
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== ops tests" ; timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -rA -p no:cacheprovider > $OUT/pytest_ops.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_ops.log
tail -5 $OUT/pytest_ops.log
echo "== forward tests" ; timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q -rA -p no:cacheprovider > $OUT/pytest_fwd.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_fwd.log
tail -5 $OUT/pytest_fwd.log
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "rc=$?" | tee -a $OUT/smoke.log
tail -3 $OUT/smoke.log
echo "== bench" ; timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"
tail -c 3000 $OUT/bench.json; tail -5 $OUT/bench.err
if [ "${1:-}" != "quick" ]; then
  echo "== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1; echo "rc=$?"
  tail -3 $OUT/ncu_bench.log
fi
