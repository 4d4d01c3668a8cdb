#!/bin/bash
# Round 2, call 14: the driver's own bench line (all legs incl. precision_tf32) + reference arm + ncu of the three pair GEMMs on the 896-wide layers
O=gpurun_out/r2c14; mkdir -p $O
timeout 900 python bench.py 2> $O/bench_w2l.err | tail -1 > $O/bench_w2l.json
python - <<'PY'
import json
l = json.load(open('gpurun_out/r2c14/bench_w2l.json'))
r = l['roofline']
print('BENCH ms %.2f e2e %.2f value %.0f conv frac %.3f burst %.3f launches %d' % (l['ms_per_step'], l['e2e']['ms_per_step'], l['value'], r['frac'], r['frac_vs_burst'], l['gpu_launches_per_step']))
print({k: (round(v['ms_per_step'], 3), round(v['frac'], 3)) for k, v in l['hbm_kernels'].items()})
print('serialized conv', r['serialized']['kernel_ms_per_step'], r['serialized']['by_pass_ms'])
for k in ('config3', 'ragged'):
    c = l.get(k, {}); print(k, c.get('ms_per_step'), c.get('bn_act_pad'), c.get('bn_act_bwd'), c.get('error'))
print('tf32', l.get('precision_tf32'))
print('loss_check', l.get('loss_check'))
print(l.get('default_config'))
print('cpu', l.get('cpu_baseline'))
PY
tail -3 $O/bench_w2l.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2> $O/bench_ref.err | tail -1 > $O/bench_reference.json
cut -c1-400 $O/bench_reference.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_cg2" -s 17 -c 1 -o $O/prof_fwd896 -f python bench.py --profile --steps 1 --warmup 0 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_cg2" -s 24 -c 1 -o $O/prof_dgrad896 -f python bench.py --profile --steps 1 --warmup 0 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_wgrad_cg2" -s 1 -c 1 -o $O/prof_wgrad896 -f python bench.py --profile --steps 1 --warmup 0 > /dev/null 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_w2l20.csv python bench.py --profile --steps 1 --warmup 1 > /dev/null 2>&1
ls -la $O
