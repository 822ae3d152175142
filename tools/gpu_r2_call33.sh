#!/bin/bash
# Round 2, call 33 (one B200): em_project v2 with batched staging loads (call 32 showed the first v2 on the long scoreboard 80 % of
# the time: one CTA per SM and element-by-element global -> shared staging), A/B against v1; full GPU tests; the default bench line.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 240 python -m pytest tests/test_gpu_ops.py tests/test_gpu_tc.py -m gpu -q -x -p no:cacheprovider -k "em_project or essential" > $OUT/pytest_guard_c33.log 2>&1; echo "guard rc=$?"; tail -3 $OUT/pytest_guard_c33.log
MAIN="--legs main --no-cpu-baseline"
timeout 300 python bench.py $MAIN > $OUT/bench_c33_new.json 2> $OUT/bench_c33_new.err; echo "bench new rc=$?"
RELPOSE_EM_PROJECT_V1=1 timeout 300 python bench.py $MAIN > $OUT/bench_c33_projv1_ab.json 2> $OUT/bench_c33_projv1_ab.err; echo "bench v1 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:em_project2_kernel -s 2 -c 1 -f -o $OUT/r2c33_em_project2_kernel python bench.py --steps 1 --warmup 1 --legs main --no-cpu-baseline --no-e2e > $OUT/ncu_em_project2.log 2>&1; echo "ncu rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_gpu_c33.log 2>&1; echo "pytest gpu rc=$?"; tail -3 $OUT/pytest_gpu_c33.log; grep -E "FAILED|Error" $OUT/pytest_gpu_c33.log | head
( time timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_c33_default.json 2> $OUT/bench_c33_default.err ) 2>&1 | grep real; echo "bench default rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c33.csv python bench.py --steps 2 --warmup 1 --legs main --no-cpu-baseline --no-e2e > $OUT/bench_under_ncu.log 2>&1; echo "ncu launch list rc=$?"
python - <<PY
import json
def last(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e: print(p, "unreadable", e); return None
for tag in ("new", "projv1_ab", "default"):
    d = last("$OUT/bench_c33_%s.json" % tag)
    if not d: continue
    print(tag, "value", round(d['value'], 1), 'e2e', d.get('e2e') and d['e2e'].get('value') and round(d['e2e']['value'], 1), 'clocks', d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
    for k, v in d['stages'].items():
        if k.startswith(('em_project', 'essential', 'mlp', 'self_att')): print(f"  {k:34s} {v['calls']:3d} {v['ms']/v['calls']*1000:8.1f} us {100*v['share']:5.1f}%")
d = last("$OUT/bench_c33_default.json")
if d:
    for k in ('roofline', 'parity', 'config4', 'config5', 'geometry', 'cpu_baseline'):
        print(k, json.dumps(d.get(k))[:400])
PY
