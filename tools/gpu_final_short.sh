#!/bin/bash
# Trimmed round-end run (fits ~4 GPU-minutes): full GPU suite, smoke, default bench, geometry microbench.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $OUT/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
tail -3 $OUT/pytest_gpu.log; grep -E "FAILED|Error" $OUT/pytest_gpu.log | head
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 300 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?"; tail -3 $OUT/bench_default.err
timeout 100 python tools/bench_geom.py > $OUT/geom.json 2> $OUT/geom.err; echo "geom rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/bench_default.json"))
print("bench_default value",round(d['value'],1),'e2e',round(d['e2e']['value'],1),'e2e_u8',round(d['e2e_u8']['value'],1),'launches',d['gpu_launches'],'clocks',d['clocks'],'cpu',d.get('cpu_baseline',{}).get('value'))
print("  roofline",{k:(round(v,4) if isinstance(v,float) else v) for k,v in d['roofline'].items() if k!='note'})
for k,v in list(d['stages'].items())[:8]: print(f"  {k:28s} {v['calls']:3d} {v['ms']:8.3f} ms {100*v['share']:5.1f}% {v['tflops']:7.2f} TF")
print({k: v["us"] for k,v in json.load(open("$OUT/geom.json"))["kernels"].items()})
PY
