#!/bin/bash
# Round 2, call 25: fp32 depthwise (separable Jasper in tf32 mode) + whole suite
O=gpurun_out/r2c25; mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_models.py -q -m gpu -p no:cacheprovider -s -k "fp32_faithful" ) 2>&1 | grep -E "tf32 mode|passed|failed" | cut -c1-260
( time timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider ) > $O/pytest_gpu.log 2>&1
grep -E "passed|failed|FAILED" $O/pytest_gpu.log | tail -5
timeout 600 python bench.py --model jasper --steps 10 --warmup 3 --skip-cpu --skip-legs --skip-default 2> $O/j.err | tail -1 | python -c "
import sys, json
l = json.loads(sys.stdin.readline())
print('separable jasper: ms_per_step %.2f  e2e %.2f' % (l['ms_per_step'], l['e2e']['ms_per_step']))"
tail -2 $O/j.err
