#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -k "essential or em_project or project or regressor_tail" > $OUT/pytest_emp.log 2>&1; echo "emp rc=$?"; tail -3 $OUT/pytest_emp.log
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q -p no:cacheprovider > $OUT/pytest_fwd.log 2>&1; echo "fwd rc=$?"; tail -2 $OUT/pytest_fwd.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_chk.json 2> $OUT/bench_chk.err; echo "bench rc=$?"; tail -2 $OUT/bench_chk.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_chk.json").read())
print("value",round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],3))
for k,v in d['stages'].items():
    if 'em_project' in k or k.startswith('linear') or 'tail' in k: print(f"  {k:32s} {v['calls']:3d} {v['ms']:8.3f} ms {100*v['share']:5.1f}%")
PY
