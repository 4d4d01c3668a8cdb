#!/bin/bash
# Round 2, call 29: kernel time budget of the literal default yaml's step (mid_layers=1, B=64 x 15 s): what is on the replayed graph's critical path
O=gpurun_out/r2c29; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_default.csv python tools/graph_vs_eager.py 1 64 15 1 > $O/ncu.log 2>&1
tail -3 $O/ncu.log
wc -l $O/launches_default.csv
