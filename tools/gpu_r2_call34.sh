#!/bin/bash
# Round 2, call 34 (one B200): regressor tail as a cluster kernel (8 CTAs per 8 rows, rows exchanged through distributed shared
# memory) against the one-row-per-CTA kernel; batched split-K reduce; pose_regressor.0 on the split-K tensor-core kernel in the
# single-plane bf16 mode too (config 4); full GPU tests; default bench line.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "regressor or em_project or linear" > $OUT/pytest_guard_c34.log 2>&1; G=$?; echo "guard rc=$G"; tail -4 $OUT/pytest_guard_c34.log
if [ $G -ne 0 ]; then grep -E "FAILED|Error|assert|parity" $OUT/pytest_guard_c34.log | head -20; echo "GUARD FAILED: the rest of the call runs with RELPOSE_REGRESSOR_TAIL_V1=1"; export RELPOSE_REGRESSOR_TAIL_V1=1; fi
MAIN="--legs main --no-cpu-baseline"
timeout 300 python bench.py $MAIN > $OUT/bench_c34_new.json 2> $OUT/bench_c34_new.err; echo "bench new rc=$?"
RELPOSE_REGRESSOR_TAIL_V1=1 timeout 300 python bench.py $MAIN > $OUT/bench_c34_tailv1_ab.json 2> $OUT/bench_c34_tailv1_ab.err; echo "bench tail v1 rc=$?"
timeout 300 python bench.py --precision bf16 --u8 $MAIN > $OUT/bench_c34_bf16_stages.json 2> $OUT/bench_c34_bf16_stages.err; echo "bench bf16 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:regressor_tail2_kernel -s 2 -c 1 -f -o $OUT/r2c34_regressor_tail2_kernel python bench.py --steps 1 --warmup 1 --legs main --no-cpu-baseline --no-e2e > $OUT/ncu_regressor_tail2.log 2>&1; echo "ncu rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_gpu_c34.log 2>&1; echo "pytest gpu rc=$?"; tail -3 $OUT/pytest_gpu_c34.log; grep -E "FAILED|Error" $OUT/pytest_gpu_c34.log | head
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_c34.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke_c34.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_c34_default.json 2> $OUT/bench_c34_default.err ) 2>&1 | grep real; echo "bench default done"
python - <<PY
import json
def last(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e: print(p, "unreadable", e); return None
for tag in ("new", "tailv1_ab", "bf16_stages", "default"):
    d = last("$OUT/bench_c34_%s.json" % tag)
    if not d: continue
    print(tag, "value", round(d['value'], 1), 'e2e', d.get('e2e') and d['e2e'].get('value') and round(d['e2e']['value'], 1), 'clocks', d['clocks'].get('sm_mhz'), d['clocks'].get('sm_min_mhz'), d['clocks'].get('reasons'))
    for k, v in d['stages'].items():
        if k.startswith(('em_project', 'regressor', 'linear_tc_splitk', 'linear[', 'layernorm', 'split_planes', 'mlp')): print(f"  {k:34s} {v['calls']:3d} {v['ms']/v['calls']*1000:8.1f} us {100*v['share']:5.1f}%")
d = last("$OUT/bench_c34_default.json")
if d:
    for k in ('roofline', 'parity', 'config4', 'config5', 'geometry', 'cpu_baseline'):
        print(k, json.dumps(d.get(k))[:400])
PY
