#!/bin/bash
# ncu evidence: launch list of one bench run + full capture of the dominant kernels (1 GPU only)
set -u
OUT=gpurun_out; mkdir -p $OUT
PREC=${1:-bf16x3}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_$PREC.csv \
    python bench.py --steps 1 --warmup 1 --batch 64 --precision $PREC --no-cpu-baseline > $OUT/ncu_list.log 2>&1; echo "list rc=$?"
for K in gemm_tc_kernel self_attention_kernel em_accum_kernel sgemm_nt_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 12 -c 2 -f -o $OUT/prof_${K}_$PREC \
      python bench.py --steps 1 --warmup 1 --batch 64 --precision $PREC --no-cpu-baseline > $OUT/ncu_$K.log 2>&1; echo "$K rc=$?"
done
ls -la $OUT/*.ncu-rep
