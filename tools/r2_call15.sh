#!/bin/bash
# Round 2, call 15 (2 GPUs): reducer parity test + driver-style N=2 bench with the CTA-pair kernels; N=1 on the same box for the ratio
O=gpurun_out/r2c15; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider ) > $O/pytest_multi.log 2>&1
tail -3 $O/pytest_multi.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 4 --skip-cpu --skip-legs --skip-default 2> $O/n1.err | tail -1 > $O/bench_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 4 2> $O/n2.err | tail -1 > $O/bench_n2.json
python - <<'PY'
import json
for n in (1, 2):
    try:
        l = json.load(open('gpurun_out/r2c15/bench_n%d.json' % n))
        print('N=%d ms %.2f e2e %.2f value %.0f reducer %s parity %s clocks %s conv_union %.2f' % (n, l['ms_per_step'], l['e2e']['ms_per_step'], l['value'], l.get('reducer'), l.get('reducer_parity'), l['clocks']['sm_mhz'], l['roofline']['kernel_ms_per_step']))
    except Exception as e:
        print(n, 'failed', e)
PY
tail -5 $O/n2.err
