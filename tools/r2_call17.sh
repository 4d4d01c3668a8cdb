#!/bin/bash
# Round 2, call 17: host profile of the small config after the host-path trims + suite
O=gpurun_out/r2c17; mkdir -p $O
timeout 300 python tools/host_profile.py 1 40 > $O/host_profile_mid1.txt 2>&1; head -40 $O/host_profile_mid1.txt
( time timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider ) > $O/pytest_gpu.log 2>&1
tail -5 $O/pytest_gpu.log
