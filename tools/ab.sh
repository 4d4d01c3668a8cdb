#!/bin/bash
# A/B of one environment knob inside ONE gpurun call (same GPU, interleaved, so clocks/thermals are comparable):
#   tools/ab.sh VAR A_VALUE B_VALUE [bench args...]
VAR=$1; A=$2; B=$3; shift 3
for rep in 1 2; do for v in $A $B; do
  env $VAR=$v timeout 300 python bench.py --steps 25 --warmup 5 --skip-default --skip-cpu --skip-legs "$@" 2>&1 | tail -1 | python -c "
import sys, json
l = json.loads(sys.stdin.readline()); r = l['roofline']
print('$VAR=$v  ms_per_step %.2f  e2e %.2f  conv_union %.2f  non_conv %.2f  clocks %s' % (l['ms_per_step'], l['e2e']['ms_per_step'], r['kernel_ms_per_step'], l['ms_per_step'] - r['kernel_ms_per_step'], l['clocks']['sm_mhz']))"
done; done
