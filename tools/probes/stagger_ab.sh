for s in 0 400 800 1200 2000 3000; do echo "stagger $s"; RELPOSE_MLP_STAGGER=$s tools/probes/mlp_trace_probe 64 2 120 | head -3; done
