#!/bin/bash
# Round 2, call 24: odd channel counts on hardware + whole suite + quick bench sanity
O=gpurun_out/r2c24; mkdir -p $O
( time timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider ) > $O/pytest_gpu.log 2>&1
grep -E "passed|failed|FAILED" $O/pytest_gpu.log | tail -5
timeout 600 python bench.py --steps 20 --warmup 4 --skip-cpu --skip-legs --skip-default 2> $O/b.err | tail -1 | python -c "
import sys, json
l = json.loads(sys.stdin.readline())
print('ms_per_step %.2f  e2e %.2f  launches %s mallocs %s  clocks %s' % (l['ms_per_step'], l['e2e']['ms_per_step'], l['gpu_launches_per_step'], l['cuda_mallocs_in_timed_region'], l['clocks']['sm_mhz']))"
