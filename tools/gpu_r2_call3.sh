#!/bin/bash
# Round 2, call 3 (one B200): Essential Matrix Module v2 + convolution accumulator folding + e2e staging ring.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -p no:cacheprovider -k "essential or conv2d_tc" > $OUT/pytest_em.log 2>&1; echo "pytest em/conv rc=$?"
tail -3 $OUT/pytest_em.log; grep -E "FAILED|Error|\[parity\]" $OUT/pytest_em.log | tail -25
timeout 1200 python -m pytest tests/test_gpu_forward.py -m gpu -q -p no:cacheprovider > $OUT/pytest_fwd.log 2>&1; echo "pytest forward rc=$?"
tail -3 $OUT/pytest_fwd.log; grep -E "FAILED|Error" $OUT/pytest_fwd.log | head; grep -E "\[parity\].*(rot_err|precision)" $OUT/pytest_fwd.log | tail -40
timeout 900 python bench.py --legs main,parity > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?"; tail -5 $OUT/bench_default.err
RELPOSE_EM_V1=1 timeout 300 python bench.py --legs main --no-cpu-baseline --no-e2e > $OUT/bench_emv1.json 2> $OUT/bench_emv1.err; echo "bench emv1 rc=$?"
RELPOSE_CONV_FOLD=0 timeout 300 python bench.py --legs main,parity --no-cpu-baseline --no-e2e > $OUT/bench_nofold.json 2> $OUT/bench_nofold.err; echo "bench nofold rc=$?"
BENCH="python bench.py --steps 1 --warmup 1 --legs main --no-cpu-baseline --no-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:em_accum2_tc_kernel -s 2 -c 1 -f -o $OUT/r2c3_em_accum2 $BENCH > $OUT/ncu_em.log 2>&1; echo "ncu em rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_pool_tc_kernel -s 2 -c 1 -f -o $OUT/r2c3_stem_pool $BENCH > $OUT/ncu_stem.log 2>&1; echo "ncu stem rc=$?"
python - <<PY
import json
for n in ("bench_default","bench_emv1","bench_nofold"):
    try:
        d=json.load(open("$OUT/%s.json"%n))
    except Exception as e:
        print(n,"unreadable",e); continue
    print(n,"value",round(d['value'],1),'e2e',d['e2e'] and round(d['e2e']['value'],1),'e2e_f32',d.get('e2e_f32') and round(d['e2e_f32']['value'],1),'launches',d['gpu_launches'],'clocks',d['clocks'])
    for k,v in list(d['stages'].items())[:14]: print(f"  {k:32s} {v['calls']:3d} {v['ms']:8.3f} ms {100*v['share']:5.1f}% {v['tflops']:7.2f} TF {v['gbs']:7.1f} GB/s")
    for k in ("parity","attention_gemm","legs_timeout"):
        if k in d: print("  ",k, json.dumps(d[k])[:1200])
PY
