#!/bin/bash
# Round 2, call 30: parallel metrics_split + metrics branch forked inside the captured step
O=gpurun_out/r2c30; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_zz_graph.py tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "graph or string_metrics or training_step" ) > $O/tests.log 2>&1
grep -E "passed|failed|FAILED|Error" $O/tests.log | tail -8 | cut -c1-300
timeout 600 python tools/graph_vs_eager.py 1 64 15 50 2>&1 | grep -v Warning | tail -3
timeout 600 python tools/graph_vs_eager.py 5 16 5 50 2>&1 | grep -v Warning | tail -3
