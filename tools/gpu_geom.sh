#!/bin/bash
# geometry kernels (config 3): parity tests + CUDA-graph microbench
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -s -p no:cacheprovider -k "se3 or svd3 or essential_to_rt or lietorch or SE3" > $OUT/pytest_geom.log 2>&1; echo "geom tests rc=$?"; tail -3 $OUT/pytest_geom.log; grep "parity\] svd3\|parity\] essential_to_rt" $OUT/pytest_geom.log
timeout 300 python tools/bench_geom.py > $OUT/geom.json 2> $OUT/geom.err; echo "geom bench rc=$?"; cat $OUT/geom.json
