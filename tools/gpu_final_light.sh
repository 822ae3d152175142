#!/bin/bash
# Round-end evidence run (one B200): full GPU test suite, smoke, default bench + reference arm + bf16 bench,
# ncu launch list and full captures of the dominant kernels, geometry microbench, training driver.
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
tail -3 $OUT/pytest_gpu.log; grep -E "FAILED|Error" $OUT/pytest_gpu.log | head
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?"; tail -3 $OUT/bench_default.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "bench ref rc=$?"
timeout 600 python bench.py --precision bf16 --no-cpu-baseline > $OUT/bench_bf16.json 2> $OUT/bench_bf16.err; echo "bench bf16 rc=$?"
timeout 300 python tools/bench_geom.py > $OUT/geom.json 2> $OUT/geom.err; echo "geom rc=$?"
timeout 600 python -m rel_pose_b200.train_synthetic --steps 20 --warmup_steps 5 --batch 6 > $OUT/train_fused.json 2> $OUT/train_fused.err; echo "train rc=$?"
BENCH="python bench.py --steps 1 --warmup 1 --batch 64 --precision bf16x3 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bf16x3.csv $BENCH > $OUT/ncu_list.log 2>&1; echo "list rc=$?"
python - <<PY
import json
for n in ("bench_default","bench_bf16"):
    d=json.load(open("$OUT/%s.json"%n))
    print(n,"value",round(d['value'],1),'e2e',round(d['e2e']['value'],1),'e2e_u8',round(d['e2e_u8']['value'],1),'launches',d['gpu_launches'],'clocks',d['clocks'],'cpu',d.get('cpu_baseline',{}).get('value'))
    print("  roofline",{k:(round(v,4) if isinstance(v,float) else v) for k,v in d['roofline'].items() if k!='note'})
    for k,v in list(d['stages'].items())[:24]: print(f"  {k:28s} {v['calls']:3d} {v['ms']:8.3f} ms {100*v['share']:5.1f}% {v['tflops']:7.2f} TF {v['gbs']:7.1f} GB/s")
print(open("$OUT/bench_reference.json").read()[:700])
print(open("$OUT/geom.json").read())
print(open("$OUT/train_fused.json").read())
PY
