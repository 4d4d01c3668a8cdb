#!/bin/bash
# Round 2, call 11: variable last N tile in the fwd-kind pair kernel: conv tests, A/B
O=gpurun_out/r2c11; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity_headline.py -q -m gpu -p no:cacheprovider -x -k "conv" ) > $O/pytest_conv.log 2>&1
tail -3 $O/pytest_conv.log
bash tools/ab.sh W2L_CG2_WIDE 0 1 2>&1 | tee $O/ab_cg2_wide.txt
ls -la $O
