#!/bin/bash
# Round 2, call 34: graph step -- version bumps after replays (stale eval caches), tf32 / padded-width modes; edit-distance loop trim
O=gpurun_out/r2c34; mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_zz_graph.py tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -s -k "graph or string_metrics" ) > $O/tests.log 2>&1
grep -E "passed|failed|FAILED|Error|^tf32|^odd_widths|^jasper eager" $O/tests.log | tail -12 | cut -c1-400
timeout 600 python tools/graph_vs_eager.py 1 64 15 50 2>&1 | grep -v Warning | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"edit_distance|metrics_" -c 5 --log-file $O/metrics_launches.csv python tools/graph_vs_eager.py 1 64 15 1 > /dev/null 2>&1
grep -E "edit_distance|metrics_" $O/metrics_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-140 | tail -5
