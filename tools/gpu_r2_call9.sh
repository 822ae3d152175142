#!/bin/bash
# Round 2, call 9 (one B200): geometry kernels after the Cauchy-Schwarz guard / first-sweep specialisation.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "svd or essential or se3" > $OUT/pytest_geom.log 2>&1; echo "pytest geom rc=$?"
tail -3 $OUT/pytest_geom.log; grep -E "FAILED|Error|assert|parity" $OUT/pytest_geom.log | head -20
timeout 300 python tools/bench_geom.py > $OUT/geom_c9.json 2> $OUT/geom_c9.err; echo "bench geom rc=$?"; tail -2 $OUT/geom_c9.err; cat $OUT/geom_c9.json | head -c 1500
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active -k regex:"svd3_kernel|essential_to_rt_kernel" -c 4 python tools/bench_geom.py > $OUT/ncu_geom.log 2>&1; grep -E "svd3_kernel|essential_to_rt|inst_executed|duration|issue_active" $OUT/ncu_geom.log | head -24
