#!/bin/bash
# Round-end evidence run on one B200 (under gpurun): tests, smoke, bench lines, yard-stick, ncu launch list + full captures.
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4
# experimental, opt-in paths (first hardware contact): the linear-domain CTC schedule (W2L_CTC_LINEAR)
W2L_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_zzz_ctc_linear.py -q -m gpu 2>&1 | tail -4
W2L_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_zzz_dw_tiled.py -q -m gpu 2>&1 | tail -4
timeout 300 python __graft_entry__.py 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 4 2>/dev/null | tail -1 > $O/final_bench_w2l.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > $O/final_bench_reference.json
timeout 600 python bench.py --model jasper10x5 --steps 8 --warmup 3 --skip-cpu 2>/dev/null | tail -1 > $O/final_bench_jasper10x5.json
timeout 600 python bench.py --model jasper --steps 10 --warmup 3 --skip-cpu 2>/dev/null | tail -1 > $O/final_bench_jasper_sep.json
# SURVEY 8d's second run: ragged lengths (Jasper's masks and the CTC length handling do real work)
timeout 600 python bench.py --ragged --steps 20 --warmup 4 --skip-cpu --skip-default 2>/dev/null | tail -1 > $O/final_bench_w2l_ragged.json
timeout 600 python bench.py --model jasper10x5 --ragged --steps 8 --warmup 3 --skip-cpu 2>/dev/null | tail -1 > $O/final_bench_jasper10x5_ragged.json
# register-tiled depthwise kernels (W2L_DW_TILED) on the shipped separable jasper.yaml: A/B on the same GPU
bash tools/ab.sh W2L_DW_TILED 0 1 --model jasper > $O/final_ab_dw_tiled.txt 2>&1
# CTC schedules side by side (same GPU, same call): log space (default) vs linear domain with / without the log-space redo
for m in 0 1 2; do W2L_CTC_LINEAR=$m timeout 600 python tools/sweep_ctc_decode.py > $O/final_ctc_sweep_lin$m.md 2>&1; done
timeout 300 python tools/yardstick_torch_cuda.py 2>&1 | tail -2 > $O/final_yardstick.txt
timeout 120 python tools/microbench_features.py 2>&1 | tail -1 > $O/final_features.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/final_launches.csv python bench.py --profile --steps 1 --warmup 1 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 16 -c 3 -o $O/final_prof_fwd -f python bench.py --profile --steps 1 --warmup 0 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 25 -c 4 -o $O/final_prof_bwd -f python bench.py --profile --steps 1 --warmup 0 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"ctc_|greedy_|bn_act|bn_finalize|novograd_|im2col" -c 16 -o $O/final_prof_misc -f python bench.py --profile --steps 1 --warmup 0 --mid-layers 1 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none -k regex:"logmel|feat_norm" -c 2 -o $O/final_prof_features -f python tools/microbench_features.py > /dev/null 2>&1
wc -c $O/final_*
cat $O/final_yardstick.txt $O/final_features.txt
