#!/bin/bash
# Round 2, call 33: warp-per-pair edit distance + inline metrics: tests, default-config replay, headline non-conv time
O=gpurun_out/r2c33; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_zz_graph.py tests/test_gpu_kernels.py tests/test_gpu_models.py -q -m gpu -p no:cacheprovider -k "graph or string_metrics or training_step" ) > $O/tests.log 2>&1
grep -E "passed|failed|FAILED|Error" $O/tests.log | tail -8 | cut -c1-300
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "string_metrics" 2>&1 | grep -E "ERROR SUMMARY|passed|failed" | tail -3
timeout 600 python tools/graph_vs_eager.py 1 64 15 50 2>&1 | grep -v Warning | tail -3
for rep in 1 2; do
  timeout 300 python bench.py --steps 25 --warmup 5 --skip-default --skip-cpu --skip-legs 2>&1 | tail -1 | python -c "
import sys, json
l = json.loads(sys.stdin.readline()); r = l['roofline']; h = l['hbm_kernels']
print('ms_per_step %.2f  e2e %.2f  conv_union %.2f  non_conv %.2f  ctc %.3f decode %.3f clocks %s' % (l['ms_per_step'], l['e2e']['ms_per_step'], r['kernel_ms_per_step'], l['ms_per_step'] - r['kernel_ms_per_step'], h['ctc_loss_raw']['ms_per_step'], h['greedy_decode']['ms_per_step'], l['clocks']['sm_mhz']))"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"edit_distance|metrics_" -c 10 --log-file $O/metrics_launches.csv python tools/graph_vs_eager.py 1 64 15 1 > /dev/null 2>&1
grep -E "edit_distance|metrics_" $O/metrics_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-140 | tail -10
