#!/bin/bash
# Round 2, call 22: cp.async-pipelined BatchNorm-backward kernels: parity, A/B in the step, ncu of both passes
O=gpurun_out/r2c22; mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py tests/test_gpu_parity_headline.py -q -m gpu -p no:cacheprovider -x ) 2>&1 | tail -3
bash tools/ab.sh W2L_BN_PIPE 0 1 2>&1 | tee $O/ab_bn_pipe_w2l.txt
for v in 0 1; do W2L_BN_PIPE=$v timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-legs --skip-default 2>/dev/null | tail -1 | python -c "
import sys, json
l = json.loads(sys.stdin.readline())
print('W2L_BN_PIPE=$v serialized:', {k: (round(v['ms_per_step'], 3), round(v['frac'], 3)) for k, v in l['hbm_kernels'].items() if k.startswith('bn')})"; done 2>&1 | tee $O/serialized_bn.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"bn_act_bwd" -s 4 -c 2 -o $O/prof_bn_bwd_pipe -f python bench.py --profile --steps 1 --warmup 0 > /dev/null 2>&1
ls -la $O
