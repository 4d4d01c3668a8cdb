#!/bin/bash
# data-parallel overlap sweep: NCCL CTAs per collective x SMs reserved for them (run under gpurun --gpus N)
N=${1:-2}
IFS=";" read -ra LIST <<< "${CFGS:-8 0;4 4;8 8}"; for cfg in "${LIST[@]}"; do
  IFS=" " read -r a b <<< "$cfg"; set -- $a $b
  echo "== N=$N NCCL_MAX_CTAS=$1 COMM_SMS=$2"
  W2L_NCCL_MAX_CTAS=$1 W2L_COMM_SMS=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29511 bench.py --gpus $N --steps ${STEPS:-20} --warmup 4 --skip-default 2>&1 | tail -1 | python -c "
import sys, json
l = json.loads(sys.stdin.readline()); r = l['roofline']
print('ms_per_step %.2f  e2e %.2f  conv_union %.2f  spans %s  clocks %s' % (l['ms_per_step'], l['e2e']['ms_per_step'], r['kernel_ms_per_step'], {k: round(v, 2) for k, v in r['by_pass_span_ms'].items()}, l['clocks']['sm_mhz']))"
done
