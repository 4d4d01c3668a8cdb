#!/bin/bash
# reducer A/B at N GPUs in one gpurun call: peer (NVLS / P2P kernel) vs NCCL
N=${1:-2}
for cfg in "peer 4" "peer 8" "nccl 4"; do
  set -- $cfg
  echo "== N=$N W2L_REDUCER=$1 ctas=$2"
  W2L_REDUCER=$1 W2L_COMM_CTAS=$2 W2L_NCCL_MAX_CTAS=$2 W2L_COMM_SMS=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-20} --warmup 4 --skip-default 2>&1 | tail -1 | python -c "
import sys, json
l = json.loads(sys.stdin.readline()); r = l['roofline']
print('ms_per_step %.2f  e2e %.2f  conv_union %.2f  clocks %s' % (l['ms_per_step'], l['e2e']['ms_per_step'], r['kernel_ms_per_step'], l['clocks']['sm_mhz']))"
done
