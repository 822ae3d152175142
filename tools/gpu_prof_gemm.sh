#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 1 --batch 64 --precision bf16x3 --no-cpu-baseline --no-e2e"
# gemm_tc launches in one forward: 13 convs (stem + 12), then per block qkv, proj, fc1, fc2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 13 -c 4 -f -o $OUT/prof_gemm2 $BENCH > $OUT/ncu_gemm2.log 2>&1; echo "gemm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 0 -c 3 -f -o $OUT/prof_conv2 $BENCH > $OUT/ncu_conv2.log 2>&1; echo "conv rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:em_ -s 0 -c 2 -f -o $OUT/prof_em2 $BENCH > $OUT/ncu_em2.log 2>&1; echo "em rc=$?"
ls -la $OUT/*2.ncu-rep
