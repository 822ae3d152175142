#!/bin/bash
# Round 2, call 4 (one B200): training step on the tensor-core engine (tests + A/B), fold retune, EM early S0.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x -p no:cacheprovider > $OUT/pytest_train.log 2>&1; echo "pytest train rc=$?"
tail -3 $OUT/pytest_train.log; grep -E "FAILED|Error|\[train\]" $OUT/pytest_train.log | tail -12
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_forward.py -m gpu -q -x -p no:cacheprovider -k "essential or golden or precisions or conv2d" > $OUT/pytest_em.log 2>&1; echo "pytest em/forward rc=$?"
tail -3 $OUT/pytest_em.log; grep -E "FAILED|Error" $OUT/pytest_em.log | head; grep -E "\[parity\].*precision=bf16x3" $OUT/pytest_em.log | tail -14
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?"; tail -5 $OUT/bench_default.err
RELPOSE_TRAIN_TC=0 timeout 300 python -m rel_pose_b200.train_synthetic --steps 10 --warmup_steps 4 > $OUT/train_simt.json 2> $OUT/train_simt.err; echo "train simt rc=$?"
timeout 300 python -m rel_pose_b200.train_synthetic --steps 10 --warmup_steps 4 > $OUT/train_tc.json 2> $OUT/train_tc.err; echo "train tc rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 1400 --csv --log-file $OUT/train_launches.csv \
    python -m rel_pose_b200.train_synthetic --steps 2 --warmup_steps 2 --batch 6 --pool 2 > $OUT/ncu_train.log 2>&1; echo "train list rc=$?"
python - <<PY
import json
for n in ("bench_default",):
    try:
        d=json.load(open("$OUT/%s.json"%n))
    except Exception as e:
        print(n,"unreadable",e); continue
    print(n,"value",round(d['value'],1),'e2e',d['e2e'] and round(d['e2e']['value'],1),'e2e_f32',d.get('e2e_f32') and round(d['e2e_f32']['value'],1),'launches',d['gpu_launches'],'clocks',d['clocks'])
    for k,v in list(d['stages'].items())[:16]: print(f"  {k:32s} {v['calls']:3d} {v['ms']:8.3f} ms {100*v['share']:5.1f}% {v['tflops']:7.2f} TF {v['gbs']:7.1f} GB/s")
    for k in ("parity","config4","config5","attention_gemm","legs_timeout","roofline"):
        if k in d: print("  ",k, json.dumps(d[k])[:1000])
for n in ("train_simt","train_tc"):
    try: print(n, open("$OUT/%s.json"%n).read()[:900])
    except Exception as e: print(n, e)
PY
