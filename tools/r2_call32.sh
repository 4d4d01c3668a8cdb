#!/bin/bash
# Round 2, call 32: where to put decode + CER/WER relative to the CTC kernels (beside them on the side stream / after them / inline)
for rep in 1 2; do for v in beside after inline; do
  env W2L_METRICS_PLACE=$v timeout 300 python bench.py --steps 25 --warmup 5 --skip-default --skip-cpu --skip-legs 2>&1 | tail -1 | python -c "
import sys, json
l = json.loads(sys.stdin.readline()); r = l['roofline']; h = l['hbm_kernels']
print('$v  ms_per_step %.2f  e2e %.2f  conv_union %.2f  non_conv %.2f  ctc %.3f decode %.3f clocks %s' % (l['ms_per_step'], l['e2e']['ms_per_step'], r['kernel_ms_per_step'], l['ms_per_step'] - r['kernel_ms_per_step'], h['ctc_loss_raw']['ms_per_step'], h['greedy_decode']['ms_per_step'], l['clocks']['sm_mhz']))"
done; done
