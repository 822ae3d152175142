#!/bin/bash
# Round 2, call 6 (one B200): LN+QKV direct epilogue + hoisted LayerNorm loads, fused MLP with the deferred output epilogue.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -p no:cacheprovider > $OUT/pytest_tc.log 2>&1; echo "pytest tc rc=$?"
tail -3 $OUT/pytest_tc.log; grep -E "FAILED|Error" $OUT/pytest_tc.log | head
timeout 1200 python -m pytest tests/test_gpu_forward.py -m gpu -q -x -p no:cacheprovider > $OUT/pytest_fwd.log 2>&1; echo "pytest forward rc=$?"
tail -3 $OUT/pytest_fwd.log; grep -E "FAILED|Error" $OUT/pytest_fwd.log | head
timeout 600 python bench.py --legs main,parity --no-cpu-baseline > $OUT/bench_c6.json 2> $OUT/bench_c6.err; echo "bench rc=$?"; tail -3 $OUT/bench_c6.err
RELPOSE_LNQKV_EPI=tma timeout 300 python bench.py --legs main --no-cpu-baseline --no-e2e > $OUT/bench_c6_tmaepi.json 2> $OUT/bench_c6_tmaepi.err; echo "bench tma-epi rc=$?"
BENCH="python bench.py --steps 1 --warmup 1 --legs main --no-cpu-baseline --no-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_fused_tc_kernel -s 2 -c 1 -f -o $OUT/r2c6_mlp $BENCH > $OUT/ncu_mlp.log 2>&1; echo "ncu mlp rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ln_linear_tc_kernel -s 2 -c 1 -f -o $OUT/r2c6_lnqkv $BENCH > $OUT/ncu_lnqkv.log 2>&1; echo "ncu lnqkv rc=$?"
python - <<PY
import json
for n in ("bench_c6","bench_c6_tmaepi"):
    try:
        d=json.load(open("$OUT/%s.json"%n))
    except Exception as e:
        print(n,"unreadable",e); continue
    print(n,"value",round(d['value'],1),'e2e',d['e2e'] and round(d['e2e']['value'],1),'launches',d['gpu_launches'],'clocks',d['clocks'])
    for k,v in list(d['stages'].items())[:16]: print(f"  {k:32s} {v['calls']:3d} {v['ms']:8.3f} ms {100*v['share']:5.1f}% {v['tflops']:7.2f} TF {v['gbs']:7.1f} GB/s")
    for k in ("parity","roofline"):
        if k in d: print("  ",k, json.dumps(d[k])[:700])
PY
