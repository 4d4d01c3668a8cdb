#!/bin/bash
# Round 2, call 41: the driver's sequence -- GPU suite, smoke, reference arm, default bench line
O=gpurun_out/r2c41; mkdir -p $O
( time timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider ) > $O/pytest_gpu.log 2>&1
tail -4 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py --impl reference ) 2> $O/ref.err | tail -1 > $O/bench_reference.json; cut -c1-300 $O/bench_reference.json; tail -4 $O/ref.err
( time timeout 900 python bench.py ) 2> $O/bench.err | tail -1 > $O/bench_w2l.json
tail -5 $O/bench.err
python - <<'PY'
import json
l = json.load(open('gpurun_out/r2c41/bench_w2l.json'))
r = l['roofline']
print('BENCH ms %.2f e2e %.2f value %.0f e2e_value %.0f conv frac %.3f burst %.3f launches %d remeasured %s mallocs %s' % (l['ms_per_step'], l['e2e']['ms_per_step'], l['value'], l['e2e']['value'], r['frac'], r['frac_vs_burst'], l['gpu_launches_per_step'], l.get('remeasured'), l['cuda_mallocs_in_timed_region']))
print({k: (round(v['ms_per_step'], 3), round(v['frac'], 3)) for k, v in l['hbm_kernels'].items()})
for k in ('config3', 'ragged', 'precision_tf32'):
    c = l.get(k, {}); print(k, c.get('ms_per_step'), c.get('error'))
print('loss_check', l.get('loss_check')); print(l.get('default_config')); print('cpu', l.get('cpu_baseline', {}).get('value'), 'traffic', r.get('traffic'))
PY
