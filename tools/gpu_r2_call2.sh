#!/bin/bash
# Round 2, call 2 (one B200): tcgen05.ld-vs-MMA probe, fused stem tests, the rest of the GPU suite, bench with the fused stem (and A/B).
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 120 tools/probes/tmem_ld_probe > $OUT/tmem_ld_probe.log 2>&1; echo "probe rc=$?"; cat $OUT/tmem_ld_probe.log
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -p no:cacheprovider -k "stem" > $OUT/pytest_stem.log 2>&1; echo "pytest stem rc=$?"
tail -3 $OUT/pytest_stem.log; grep -E "FAILED|Error|\[parity\]|\[stem\]" $OUT/pytest_stem.log | head -20
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
tail -5 $OUT/pytest_gpu.log; grep -E "FAILED|Error" $OUT/pytest_gpu.log | head -30
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?"; tail -5 $OUT/bench_default.err
RELPOSE_FUSED_STEM=0 timeout 300 python bench.py --legs main --no-cpu-baseline --no-e2e > $OUT/bench_stem0.json 2> $OUT/bench_stem0.err; echo "bench stem0 rc=$?"
python - <<PY
import json
for n in ("bench_default","bench_stem0"):
    try:
        d=json.load(open("$OUT/%s.json"%n))
    except Exception as e:
        print(n,"unreadable",e); continue
    print(n,"value",round(d['value'],1),'e2e',d['e2e'] and round(d['e2e']['value'],1),'e2e_f32',d.get('e2e_f32') and round(d['e2e_f32']['value'],1),'launches',d['gpu_launches'],'clocks',d['clocks'])
    for k,v in list(d['stages'].items())[:22]: print(f"  {k:32s} {v['calls']:3d} {v['ms']:8.3f} ms {100*v['share']:5.1f}% {v['tflops']:7.2f} TF {v['gbs']:7.1f} GB/s")
    for k in ("parity","gpu_eager_baseline","config4","config5","geometry","cpu_baseline","attention_gemm","legs_timeout"):
        if k in d: print("  ",k, json.dumps(d[k])[:900])
PY
