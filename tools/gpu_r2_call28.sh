#!/bin/bash
# Round 2, call 28 (one B200): fused MLP with the in-place output warpgroup (vector reductions): timeline probe, parity tests, bench A/B.
set -u
OUT=gpurun_out; mkdir -p $OUT
tools/probes/mlp_trace_probe 64 2 120 0 > $OUT/mlp_trace_outofplace.txt 2>&1; head -3 $OUT/mlp_trace_outofplace.txt
tools/probes/mlp_trace_probe 64 2 120 1 > $OUT/mlp_trace_inplace.txt 2>&1; head -3 $OUT/mlp_trace_inplace.txt
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -p no:cacheprovider -x -k "mlp" > $OUT/pytest_tc_mlp.log 2>&1; echo "pytest mlp rc=$?"; tail -3 $OUT/pytest_tc_mlp.log
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q -p no:cacheprovider -x > $OUT/pytest_fwd.log 2>&1; echo "pytest forward rc=$?"; tail -3 $OUT/pytest_fwd.log
timeout 600 python bench.py --steps 20 --warmup 5 --legs main,parity --no-cpu-baseline > $OUT/bench_c28_inplace.json 2> $OUT/bench_c28_inplace.err; echo "bench rc=$?"; head -c 500 $OUT/bench_c28_inplace.json; echo
RELPOSE_MLP_INPLACE=0 timeout 600 python bench.py --steps 20 --warmup 5 --legs main --no-cpu-baseline > $OUT/bench_c28_outofplace_ab.json 2> $OUT/bench_c28_outofplace_ab.err; echo "bench A/B rc=$?"; head -c 500 $OUT/bench_c28_outofplace_ab.json; echo
