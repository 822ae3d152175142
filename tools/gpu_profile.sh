#!/bin/bash
# ncu evidence: launch list of one bench run + full capture of the dominant kernels (1 GPU only)
set -u
OUT=gpurun_out; mkdir -p $OUT
PREC=${1:-bf16x3}
BENCH="python bench.py --steps 1 --warmup 1 --batch 64 --precision $PREC --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$PREC.csv \
    $BENCH > $OUT/ncu_list.log 2>&1; echo "list rc=$?"
# transformer GEMMs of block 0 (the 12 launches before them are the CNN convolutions): qkv, proj, fc1, fc2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 12 -c 4 -f -o $OUT/prof_gemm_$PREC \
    $BENCH > $OUT/ncu_gemm.log 2>&1; echo "gemm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 0 -c 12 -f -o $OUT/prof_conv_$PREC \
    $BENCH > $OUT/ncu_conv.log 2>&1; echo "conv rc=$?"
for K in self_attention_tc_kernel em_accum_kernel em_stats_kernel conv2d_nhwc; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 0 -c 1 -f -o $OUT/prof_${K}_$PREC \
      $BENCH > $OUT/ncu_$K.log 2>&1; echo "$K rc=$?"
done
ls -la $OUT/*.ncu-rep
