#!/bin/bash
# Round 2, call 27: GraphedTrainStep (whole-step CUDA graph) tests + the default-config leg of the bench
O=gpurun_out/r2c27; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_zz_graph.py -q -m gpu -p no:cacheprovider -s ) > $O/graph_tests.log 2>&1
tail -40 $O/graph_tests.log | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 --skip-cpu --skip-legs 2> $O/b.err | tail -1 > $O/bench.json
python - <<'P'
import json
l = json.loads(open('gpurun_out/r2c27/bench.json').read())
print('w2l20 ms/step', l['ms_per_step'], 'e2e', l['e2e']['ms_per_step'])
print(json.dumps(l.get('default_config'), indent=1))
P
tail -3 $O/b.err
