#!/bin/bash
# Round 2, call 36: Jasper 10x5 per-layer GEMM table (how much time sits in the short-K 1x1 residual GEMMs?)
O=gpurun_out/r2c36; mkdir -p $O
timeout 900 python bench.py --model jasper10x5 --steps 8 --warmup 3 --skip-cpu --skip-legs --skip-default 2> $O/j.err | tail -1 > $O/bench_jasper10x5.json
python - <<'P'
import json
l = json.load(open('gpurun_out/r2c36/bench_jasper10x5.json'))
print('ms', l['ms_per_step'], 'conv union', l['roofline']['kernel_ms_per_step'])
s = l['roofline']['serialized']
print('serialized', s['kernel_ms_per_step'], s['by_pass_ms'])
rows = sorted(s['by_layer'], key=lambda r: -r['ms_per_step'])
k1 = sum(r['ms_per_step'] for r in rows if r['k'] == 1)
print('k=1 GEMMs total ms', k1)
for r in rows:
    print('%-16s %4d->%4d k=%2d d=%d calls %.0f  %.3f ms  %.0f TF/s' % (r['pass'], r['Cin'], r['Cout'], r['k'], r['dilation'], r['calls_per_step'], r['ms_per_step'], r['tflops']))
P
tail -2 $O/j.err
