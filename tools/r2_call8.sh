#!/bin/bash
# Round 2, call 8: first hardware contact of the CTA-pair GEMM (cta_group::2) + the BatchNorm-backward load-phase fix.
O=gpurun_out/r2c8; mkdir -p $O
# the conv kernel tests first (a protocol bug traps after 2^26 polls instead of hanging)
( timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -x -k "conv" ) > $O/pytest_conv.log 2>&1
tail -5 $O/pytest_conv.log
( time timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider ) > $O/pytest_gpu.log 2>&1
tail -6 $O/pytest_gpu.log
bash tools/ab.sh W2L_CG2 0 1 2>&1 | tee $O/ab_cg2.txt
timeout 900 python bench.py --steps 20 --warmup 4 --skip-cpu 2> $O/bench_w2l.err | tail -1 > $O/bench_w2l.json
python - <<'PY'
import json
l = json.load(open('gpurun_out/r2c8/bench_w2l.json'))
r = l['roofline']
print('BENCH ms %.2f e2e %.2f value %.0f conv frac %.3f burst %.3f launches %d' % (l['ms_per_step'], l['e2e']['ms_per_step'], l['value'], r['frac'], r['frac_vs_burst'], l['gpu_launches_per_step']))
print({k: (round(v['ms_per_step'], 3), round(v['frac'], 3)) for k, v in l['hbm_kernels'].items()})
print('serialized conv', r['serialized']['kernel_ms_per_step'], r['serialized']['by_pass_ms'])
for k in ('config3', 'ragged'):
    c = l.get(k, {}); print(k, c.get('ms_per_step'), c.get('bn_act_pad'), c.get('bn_act_bwd'))
print(l.get('default_config'))
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_cg2" -s 16 -c 2 -o $O/prof_cg2_fwd -f python bench.py --profile --steps 1 --warmup 0 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"bn_act_bwd" -s 4 -c 2 -o $O/prof_bn_bwd -f python bench.py --profile --steps 1 --warmup 0 > /dev/null 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_w2l20.csv python bench.py --profile --steps 1 --warmup 1 > /dev/null 2>&1
ls -la $O
