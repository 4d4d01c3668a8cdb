#!/bin/bash
# ncu on the BatchNorm passes only (small captures: the merge back is limited to 64 MiB) + launch list
O=gpurun_out/r2c7; mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"bn_act_bwd" -s 4 -c 4 -o $O/prof_bn_bwd -f python bench.py --profile --steps 1 --warmup 0 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"bn_act_pad" -s 17 -c 2 -o $O/prof_bn_fwd -f python bench.py --profile --steps 1 --warmup 0 > /dev/null 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_w2l20.csv python bench.py --profile --steps 1 --warmup 1 > /dev/null 2>&1
ls -la $O
