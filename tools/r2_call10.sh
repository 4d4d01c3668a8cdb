#!/bin/bash
# Round 2, call 10: CTA-pair wgrad with flattened (ci tile, tap) pairing and a narrow last N tile: conv tests, A/B, ncu
O=gpurun_out/r2c10; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity_headline.py -q -m gpu -p no:cacheprovider -x -k "conv" ) > $O/pytest_conv.log 2>&1
tail -3 $O/pytest_conv.log
bash tools/ab.sh W2L_CG2_WGRAD 0 1 2>&1 | tee $O/ab_cg2_wgrad.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_wgrad_cg2" -s 0 -c 4 -o $O/prof_cg2_wgrad -f python bench.py --profile --steps 1 --warmup 0 > /dev/null 2>&1
ls -la $O
