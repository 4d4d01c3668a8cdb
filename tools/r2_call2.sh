#!/bin/bash
# Round 2, call 2: GPU suite after the test restructuring, host-time profile of the launch-bound default config (6.0 ms vs 2.1 ms in round 1?)
O=gpurun_out/r2c2
mkdir -p $O
( time timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider ) > $O/pytest_gpu.log 2>&1
tail -4 $O/pytest_gpu.log
timeout 300 python tools/host_profile.py 1 40 > $O/host_profile_mid1.txt 2>&1; head -60 $O/host_profile_mid1.txt
W2L_WGRAD_STREAM=0 timeout 300 python tools/host_profile.py 1 40 2>&1 | head -3
timeout 300 python bench.py --mid-layers 1 --steps 20 --warmup 5 --skip-cpu --skip-legs --skip-default 2>/dev/null | tail -1 | python -c "
import sys, json
l = json.loads(sys.stdin.readline()); print('mid1 bench ms', l['ms_per_step'], 'e2e', l['e2e']['ms_per_step'], 'launches', l['gpu_launches_per_step'], l['step_ms_min_median_max'])"
