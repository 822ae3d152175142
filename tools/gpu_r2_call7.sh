#!/bin/bash
# Round 2, call 7 (one B200): LayerNorms moved to the producers' epilogues (plane chain), programmatic dependent launch.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -p no:cacheprovider > $OUT/pytest_tc.log 2>&1; echo "pytest tc rc=$?"
tail -3 $OUT/pytest_tc.log; grep -E "FAILED|Error|assert" $OUT/pytest_tc.log | head -20
timeout 1200 python -m pytest tests/test_gpu_forward.py tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider > $OUT/pytest_fwd.log 2>&1; echo "pytest forward/ops rc=$?"
tail -3 $OUT/pytest_fwd.log; grep -E "FAILED|Error" $OUT/pytest_fwd.log | head
timeout 600 python bench.py --legs main,parity --no-cpu-baseline > $OUT/bench_c7.json 2> $OUT/bench_c7.err; echo "bench rc=$?"; tail -3 $OUT/bench_c7.err
RELPOSE_CHAIN_LN=0 timeout 300 python bench.py --legs main --no-cpu-baseline --no-e2e > $OUT/bench_c7_nochain.json 2> $OUT/bench_c7_nochain.err; echo "bench nochain rc=$?"
RELPOSE_PDL=0 timeout 300 python bench.py --legs main --no-cpu-baseline --no-e2e > $OUT/bench_c7_nopdl.json 2> $OUT/bench_c7_nopdl.err; echo "bench nopdl rc=$?"
BENCH="python bench.py --steps 1 --warmup 1 --legs main --no-cpu-baseline --no-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_fused_tc_kernel -s 2 -c 1 -f -o $OUT/r2c7_mlp $BENCH > $OUT/ncu_mlp.log 2>&1; echo "ncu mlp rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ln_linear_tc_kernel -s 2 -c 2 -f -o $OUT/r2c7_lnlin $BENCH > $OUT/ncu_lnlin.log 2>&1; echo "ncu lnlin rc=$?"
python - <<PY
import json
for n in ("bench_c7","bench_c7_nochain","bench_c7_nopdl"):
    try:
        d=json.load(open("$OUT/%s.json"%n))
    except Exception as e:
        print(n,"unreadable",e); continue
    print(n,"value",round(d['value'],1),'e2e',d['e2e'] and round(d['e2e']['value'],1),'launches',d['gpu_launches'],'clocks',d['clocks'])
    for k,v in list(d['stages'].items())[:18]: print(f"  {k:32s} {v['calls']:3d} {v['ms']:8.3f} ms {100*v['share']:5.1f}% {v['tflops']:7.2f} TF {v['gbs']:7.1f} GB/s")
    for k in ("parity","roofline"):
        if k in d: print("  ",k, json.dumps(d[k])[:600])
PY
