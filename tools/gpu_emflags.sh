#!/bin/bash
# Essential Matrix Module ablation flags on the tcgen05 kernels: op tests, whole-path parity, short bench
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -s -p no:cacheprovider -k "essential" > $OUT/pytest_em_tc.log 2>&1; echo "tc rc=$?"; tail -2 $OUT/pytest_em_tc.log; grep "parity\] essential_tc flags" $OUT/pytest_em_tc.log
timeout 300 python -m pytest tests/test_gpu_forward.py -m gpu -q -s -p no:cacheprovider > $OUT/pytest_em_fwd.log 2>&1; echo "fwd rc=$?"; tail -2 $OUT/pytest_em_fwd.log; grep "parity\] ablate_.*precision=bf16x3\|parity\] ablate_[a-z_0-9]*: rot" $OUT/pytest_em_fwd.log
grep -E "FAILED|Error" $OUT/pytest_em_tc.log $OUT/pytest_em_fwd.log | head
timeout 200 python tools/bench_ablations.py > $OUT/ablations.json 2> $OUT/ablations.err; echo "abl rc=$?"; cat $OUT/ablations.json
