#!/bin/bash
# Round 2, call 38: separate BatchNorm-statistics pass for the k = 1 convs (short mainloops) vs statistics in their GEMM epilogue
bash tools/ab.sh W2L_STATS_PASS_K1 0 1 --model jasper10x5 --steps 10
bash tools/ab.sh W2L_STATS_PASS_K1 0 1
( timeout 900 python -m pytest tests/test_gpu_models.py -q -m gpu -p no:cacheprovider -x ) 2>&1 | tail -2
