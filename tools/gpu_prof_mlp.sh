#!/bin/bash
# ncu full captures (source-level) of the fused MLP and attention kernels, one launch each
set -u
OUT=gpurun_out; mkdir -p $OUT
PREC=${1:-bf16x3}
BENCH="python bench.py --steps 1 --warmup 1 --batch 64 --precision $PREC --no-cpu-baseline --no-e2e"
for K in ${2:-mlp_fused_tc_kernel self_attention_tc_kernel}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o $OUT/prof_${K}_$PREC \
      $BENCH > $OUT/ncu_$K.log 2>&1; echo "$K rc=$?"; tail -2 $OUT/ncu_$K.log
done
ls -la $OUT/*.ncu-rep
