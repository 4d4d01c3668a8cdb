#!/bin/bash
# Round 2, call 42: BatchNorm-backward occupancy A/B -- 2 CTAs/SM x 4 rows in flight (default, 128 registers) vs 3 CTAs/SM (80 registers,
# spills) with 4 or 2 rows in flight.  The library is rebuilt on the box for every variant (the box copy is scratch).
SRC=wav2letter_pytorch_b200/csrc/elementwise.cu
cp $SRC /tmp/elementwise.orig.cu
variant() {
  cp /tmp/elementwise.orig.cu $SRC
  case $1 in
    v1) sed -i 's/constexpr int kBnBwdCtasPerSm = 2;/constexpr int kBnBwdCtasPerSm = 3;/' $SRC ;;
    v2) sed -i -e 's/constexpr int kBnBwdCtasPerSm = 2;/constexpr int kBnBwdCtasPerSm = 3;/' -e 's/constexpr int kRows = (HAS_RES || sizeof(TA) == 4) ? 2 : kRowGroup;/constexpr int kRows = 2;/g' $SRC ;;
  esac
  python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
}
for rep in 1 2; do for v in base v1 v2; do
  variant $v
  timeout 300 python bench.py --steps 25 --warmup 5 --skip-default --skip-cpu --skip-legs 2>&1 | tail -1 | python -c "
import sys, json
l = json.loads(sys.stdin.readline()); r = l['roofline']; h = l['hbm_kernels']
print('$v  ms_per_step %.2f  e2e %.2f  conv_union %.2f  non_conv %.2f  bn_bwd %.3f (%.3f of hbm)  clocks %s' % (l['ms_per_step'], l['e2e']['ms_per_step'], r['kernel_ms_per_step'], l['ms_per_step'] - r['kernel_ms_per_step'], h['bn_act_bwd']['ms_per_step'], h['bn_act_bwd']['frac'], l['clocks']['sm_mhz']))"
done; done
for v in base v2; do
  variant $v
  timeout 300 python bench.py --model jasper10x5 --steps 10 --warmup 3 --skip-default --skip-cpu --skip-legs 2>&1 | tail -1 | python -c "
import sys, json
l = json.loads(sys.stdin.readline()); r = l['roofline']; h = l['hbm_kernels']
print('jasper10x5 $v  ms_per_step %.2f  conv_union %.2f  non_conv %.2f  bn_bwd %.3f (%.3f of hbm)' % (l['ms_per_step'], r['kernel_ms_per_step'], l['ms_per_step'] - r['kernel_ms_per_step'], h['bn_act_bwd']['ms_per_step'], h['bn_act_bwd']['frac']))"
done
cp /tmp/elementwise.orig.cu $SRC
