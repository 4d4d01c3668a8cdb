#!/bin/bash
# Round 2, call 44: ncu launch list of the final state (W2L-20, one settled step) -- the per-kernel shares behind the bench line
O=gpurun_out/r2c44; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_w2l20.csv python bench.py --profile --steps 1 --warmup 1 > /dev/null 2>&1
wc -l $O/launches_w2l20.csv
