#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 1 --batch 64 --precision bf16x3 --no-cpu-baseline --no-e2e"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:self_attention_tc -s 0 -c 1 -f -o $OUT/prof_att3 $BENCH > $OUT/ncu_att3.log 2>&1; echo "att rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 1 -c 2 -f -o $OUT/prof_conv3 $BENCH > $OUT/ncu_conv3.log 2>&1; echo "conv rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 15 -c 2 -f -o $OUT/prof_gemm3 $BENCH > $OUT/ncu_gemm3.log 2>&1; echo "gemm rc=$?"
