#!/bin/bash
# Round 2, call 35 (8 GPUs): driver-style scaling run N = 1, 2, 4, 8 at HEAD (metrics inline behind the CTC kernels) + reducer parity test
O=gpurun_out/r2c35; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider ) > $O/pytest_multi.log 2>&1
tail -2 $O/pytest_multi.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 4 --skip-cpu --skip-legs --skip-default 2> $O/n1.err | tail -1 > $O/bench_n1.json
for n in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 4 2> $O/n$n.err | tail -1 > $O/bench_n$n.json
done
python - <<'PY'
import json
base = None
for n in (1, 2, 4, 8):
    try:
        l = json.load(open('gpurun_out/r2c35/bench_n%d.json' % n))
        base = base or l['value']
        print('N=%d ms %.2f e2e %.2f value %.0f eff %.3f reducer %s parity %s clocks %s conv_union %.2f' % (n, l['ms_per_step'], l['e2e']['ms_per_step'], l['value'], l['value'] / (n * base), l.get('reducer'), l.get('reducer_parity'), l['clocks']['sm_mhz'], l['roofline']['kernel_ms_per_step']))
    except Exception as e:
        print(n, 'failed', e)
PY
tail -3 $O/n8.err
