#!/bin/bash
# Round 2, call 9: CTA-pair wgrad (tap pairs share the dy tile) on hardware: conv tests, suite, A/B against the single-CTA wgrad
O=gpurun_out/r2c9; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity_headline.py -q -m gpu -p no:cacheprovider -x -k "conv" ) > $O/pytest_conv.log 2>&1
tail -5 $O/pytest_conv.log
( time timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider ) > $O/pytest_gpu.log 2>&1
tail -6 $O/pytest_gpu.log
bash tools/ab.sh W2L_CG2_WGRAD 0 1 2>&1 | tee $O/ab_cg2_wgrad.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_wgrad_cg2" -s 2 -c 2 -o $O/prof_cg2_wgrad -f python bench.py --profile --steps 1 --warmup 0 > /dev/null 2>&1
ls -la $O
