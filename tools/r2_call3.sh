#!/bin/bash
O=gpurun_out/r2c3; mkdir -p $O
timeout 300 python tools/host_profile.py 1 40 --after-big > $O/hp_after_big.txt 2>&1; head -40 $O/hp_after_big.txt
timeout 300 python tools/host_profile.py 1 40 --after-big --no-empty-cache 2>&1 | head -2
