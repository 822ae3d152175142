#!/bin/bash
# Round 2, call 1 (one B200): full GPU test suite, default bench with every leg, attention one-vs-two CTAs per SM A/B,
# ncu launch list of the training step and a full capture of the attention kernel.
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
tail -5 $OUT/pytest_gpu.log; grep -E "FAILED|Error|\[parity\]" $OUT/pytest_gpu.log | head -30
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?"; tail -5 $OUT/bench_default.err
RELPOSE_ATT_CPS=1 timeout 300 python bench.py --legs main --no-cpu-baseline --no-e2e > $OUT/bench_att_cps1.json 2> $OUT/bench_att_cps1.err; echo "bench cps1 rc=$?"
timeout 300 python bench.py --legs main --no-cpu-baseline --no-e2e > $OUT/bench_att_cps2.json 2> $OUT/bench_att_cps2.err; echo "bench cps2 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 1600 --csv --log-file $OUT/train_launches.csv \
    python -m rel_pose_b200.train_synthetic --steps 2 --warmup_steps 2 --batch 6 --pool 2 > $OUT/ncu_train.log 2>&1; echo "train list rc=$?"
BENCH="python bench.py --steps 1 --warmup 1 --legs main --no-cpu-baseline --no-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:self_attention_tc_kernel -s 2 -c 1 -f -o $OUT/r2c1_self_attention $BENCH > $OUT/ncu_att.log 2>&1; echo "ncu att rc=$?"
python - <<PY
import json
for n in ("bench_default","bench_att_cps1","bench_att_cps2"):
    try:
        d=json.load(open("$OUT/%s.json"%n))
    except Exception as e:
        print(n,"unreadable",e); continue
    print(n,"value",round(d['value'],1),'e2e',d['e2e'] and round(d['e2e']['value'],1),'launches',d['gpu_launches'],'clocks',d['clocks'])
    for k,v in list(d['stages'].items())[:22]: print(f"  {k:32s} {v['calls']:3d} {v['ms']:8.3f} ms {100*v['share']:5.1f}% {v['tflops']:7.2f} TF {v['gbs']:7.1f} GB/s")
    for k in ("parity","gpu_eager_baseline","config4","config5","geometry","cpu_baseline","attention_gemm","legs_timeout"):
        if k in d: print("  ",k, json.dumps(d[k])[:1500])
PY
