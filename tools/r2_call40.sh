#!/bin/bash
# Round 2, call 40: Jasper at internally padded channel counts (jasper_odd fixture) + the model / strided / graph files
O=gpurun_out/r2c40; mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_zz_strided.py tests/test_gpu_zz_graph.py -q -m gpu -p no:cacheprovider ) > $O/tests.log 2>&1
grep -E "passed|failed|FAILED|Error" $O/tests.log | tail -8 | cut -c1-300
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_models.py -q -m gpu -p no:cacheprovider -k "jasper_odd" 2>&1 | grep -E "ERROR SUMMARY|passed|failed" | tail -3
timeout 600 python bench.py --model jasper --steps 10 --warmup 3 --skip-cpu --skip-legs --skip-default 2> $O/j.err | tail -1 | python -c "
import sys, json
l = json.loads(sys.stdin.readline())
print('separable jasper: ms_per_step %.2f  e2e %.2f' % (l['ms_per_step'], l['e2e']['ms_per_step']))"
