#!/bin/bash
# Round 2, call 38 (one B200): attention projection in place (x += a W^T + b as 16-byte reductions at the memory side, no
# transposing patch / residual loads) against the load-add-store epilogue (RELPOSE_LINEAR_RED=0): bit-identity tests, A/B bench.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_forward.py -m gpu -q -x -p no:cacheprovider > $OUT/pytest_c38.log 2>&1; echo "pytest tc+forward rc=$?"; tail -3 $OUT/pytest_c38.log; grep -E "FAILED|Error" $OUT/pytest_c38.log | head
MAIN="--legs main --no-cpu-baseline"
timeout 200 python bench.py $MAIN > $OUT/bench_c38_new.json 2> $OUT/bench_c38_new.err; echo "bench new rc=$?"
RELPOSE_LINEAR_RED=0 timeout 200 python bench.py $MAIN > $OUT/bench_c38_red_off_ab.json 2> $OUT/bench_c38_red_off_ab.err; echo "bench red off rc=$?"
timeout 200 python bench.py $MAIN > $OUT/bench_c38_new2.json 2> $OUT/bench_c38_new2.err; echo "bench new repeat rc=$?"
python - <<PY
import json
def last(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e: print(p, "unreadable", e); return None
for tag in ("new", "red_off_ab", "new2"):
    d = last("$OUT/bench_c38_%s.json" % tag)
    if not d: continue
    print(tag, "value", round(d['value'], 1), 'e2e', d.get('e2e') and d['e2e'].get('value') and round(d['e2e']['value'], 1), 'clocks', d['clocks'].get('sm_mhz'), d['clocks'].get('sm_min_mhz'), d['clocks'].get('reasons'))
    for k, v in d['stages'].items():
        if k.startswith(('linear_tc', 'mlp', 'self_att', 'ln_linear')): print(f"  {k:34s} {v['calls']:3d} {v['ms']/v['calls']*1000:8.1f} us {100*v['share']:5.1f}%")
PY
