#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -rA -p no:cacheprovider -k "essential_tc or self_attention" > $OUT/pytest_em2.log 2>&1; echo "rc=$?"
grep -E "parity\] (em_tc|self)|tc-diag|FAILED|passed|failed" $OUT/pytest_em2.log | cut -c1-420 | head -40
timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -p no:cacheprovider -k "essential_tc and 30-1.0-True-1" > $OUT/san.log 2>&1
echo "== sanitizer P=1 B=30: rc=$? $(grep 'ERROR SUMMARY' $OUT/san.log | head -1)"; grep -o "essential_tc.cu:[0-9]*" $OUT/san.log | sort | uniq -c | head -5
