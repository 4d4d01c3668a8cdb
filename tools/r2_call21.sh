#!/bin/bash
# Round 2, call 21: e2e leg through DevicePrefetcher (H2D of the next batch beside the current step), prefetcher + fused-reduce model tests
O=gpurun_out/r2c21; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_zz_loader.py tests/test_gpu_models.py -q -m gpu -p no:cacheprovider -k "prefetcher or fused_bn_reduce" ) 2>&1 | tail -3
for rep in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 4 --skip-cpu --skip-legs --skip-default 2> $O/b.err | tail -1 | python -c "
import sys, json
l = json.loads(sys.stdin.readline())
print('ms_per_step %.2f  e2e %.2f  mallocs %s  clocks %s' % (l['ms_per_step'], l['e2e']['ms_per_step'], l['cuda_mallocs_in_timed_region'], l['clocks']['sm_mhz']))"
done
tail -3 $O/b.err
