#!/bin/bash
# Round 2, first GPU call: the whole GPU suite without -x (first hardware contact of everything behind round 1's red test and of the
# new headline-parity tests), the opt-in kernel families with their A/Bs, compute-sanitizer passes, smoke, the bench line, launch list.
O=gpurun_out/r2c1
mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv,noheader > $O/gpu.txt
( time timeout 1500 python -m pytest tests -q -m gpu -s -p no:cacheprovider ) > $O/pytest_gpu.log 2>&1
tail -5 $O/pytest_gpu.log
W2L_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_zzz_ctc_linear.py tests/test_gpu_zzz_dw_tiled.py -q -m gpu -p no:cacheprovider > $O/pytest_experimental.log 2>&1
tail -3 $O/pytest_experimental.log
timeout 300 python __graft_entry__.py > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 4 2> $O/bench_w2l.err | tail -1 > $O/bench_w2l.json
python - <<'PY'
import json
l = json.load(open('gpurun_out/r2c1/bench_w2l.json'))
r = l['roofline']
print('BENCH ms %.2f e2e %.2f value %.0f conv frac %.3f burst %.3f launches %d' % (l['ms_per_step'], l['e2e']['ms_per_step'], l['value'], r['frac'], r['frac_vs_burst'], l['gpu_launches_per_step']))
print({k: (round(v['ms_per_step'], 3), round(v['frac'], 3)) for k, v in l['hbm_kernels'].items()})
for k in ('config3', 'ragged', 'loss_check'):
    print(k, json.dumps(l.get(k))[:600])
for row in l.get('config5', {}).get('rows', []):
    print(row)
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > $O/bench_reference.json
for m in 0 1 2; do W2L_CTC_LINEAR=$m timeout 300 python tools/sweep_ctc_decode.py --quick > $O/ctc_sweep_lin$m.md 2>&1; done
tail -n 8 $O/ctc_sweep_lin*.md
bash tools/ab.sh W2L_DW_TILED 0 1 --model jasper > $O/ab_dw_tiled.txt 2>&1; cat $O/ab_dw_tiled.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_w2l20.csv python bench.py --profile --steps 1 --warmup 1 > /dev/null 2>&1
wc -l $O/launches_w2l20.csv
# compute-sanitizer (caching allocator off so that every tensor is its own allocation)
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
timeout 600 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_zz_strided.py -q -m gpu -p no:cacheprovider \
   -k "not full_size and not 64-750 and not slab and not config1 and not golden_and_ragged" > $O/sanitizer_memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $O/sanitizer_memcheck.log | tail -8
timeout 420 compute-sanitizer --tool initcheck --print-limit 30 python -m pytest tests/test_gpu_zz_strided.py tests/test_gpu_models.py -q -m gpu -p no:cacheprovider \
   -k "w2l_golden or training_step or im2col_tm or depthwise_dgrad or unfold or head_block or jasper_dense" > $O/sanitizer_initcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed|Uninitialized" $O/sanitizer_initcheck.log | tail -8
