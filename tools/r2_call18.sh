#!/bin/bash
# Round 2, call 18 (8 GPUs): weak and strong scaling of the W2L-20 step with the CTA-pair kernels; N=1 on the same box for the ratio
O=gpurun_out/r2c18; mkdir -p $O
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 4 --skip-cpu --skip-legs --skip-default 2> $O/n1.err | tail -1 > $O/bench_n1.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 4 2> $O/n8.err | tail -1 > $O/bench_n8.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 4 --strong 2> $O/n8s.err | tail -1 > $O/bench_n8_strong.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 4 2> $O/n4.err | tail -1 > $O/bench_n4.json
python - <<'PY'
import json
for tag in ('n1', 'n4', 'n8', 'n8_strong'):
    try:
        l = json.load(open('gpurun_out/r2c18/bench_%s.json' % tag))
        print('%s: n_gpus %d scaling %s ms %.2f e2e %.2f value %.0f reducer %s parity %s clocks %s conv_union %.2f batch %s' % (tag, l['n_gpus'], l['scaling'], l['ms_per_step'], l['e2e']['ms_per_step'], l['value'], l.get('reducer'), str(l.get('reducer_parity'))[:40], l['clocks']['sm_mhz'], l['roofline']['kernel_ms_per_step'], l['config']['global_batch']))
    except Exception as e:
        print(tag, 'failed', e)
PY
tail -3 $O/n8.err $O/n8s.err
