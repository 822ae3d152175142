#!/bin/bash
# Full GPU visit: every -m gpu test, smoke, default bench (+ reference arm), both precisions.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
tail -4 $OUT/pytest_gpu.log; grep -E "drop-in|FAILED|Error" $OUT/pytest_gpu.log | head
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?"; tail -3 $OUT/bench_default.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "bench ref rc=$?"; cat $OUT/bench_reference.json | cut -c1-600
timeout 600 python bench.py --precision bf16 --no-cpu-baseline > $OUT/bench_bf16.json 2> $OUT/bench_bf16.err; echo "bench bf16 rc=$?"
python - <<PY
import json
for n in ("bench_default","bench_bf16"):
    d=json.load(open("$OUT/%s.json"%n))
    print(n,"value",round(d['value'],1),'e2e',round(d['e2e']['value'],1),'e2e_u8',round(d['e2e_u8']['value'],1),'launches',d['gpu_launches'],'clocks',d['clocks'],'cpu',d.get('cpu_baseline',{}).get('value'))
    print("  roofline",{k:(round(v,4) if isinstance(v,float) else v) for k,v in d['roofline'].items() if k!='note'})
    for k,v in list(d['stages'].items())[:12]: print(f"  {k:28s} {v['calls']:3d} {v['ms']:8.3f} ms {100*v['share']:5.1f}% {v['tflops']:7.2f} TF {v['gbs']:7.1f} GB/s")
PY
