#!/bin/bash
# Round 2, call 20: the BatchNorm-backward reduction folded into the backward-data GEMM epilogue: kernel test, suite, A/B (W2L-20 and Jasper 10x5)
O=gpurun_out/r2c20; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -x -k "fused_bn_reduce" ) 2>&1 | tail -3
( time timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider ) > $O/pytest_gpu.log 2>&1
tail -5 $O/pytest_gpu.log
bash tools/ab.sh W2L_FUSE_BN_REDUCE 0 1 2>&1 | tee $O/ab_fuse_w2l.txt
bash tools/ab.sh W2L_FUSE_BN_REDUCE 0 1 --model jasper10x5 2>&1 | tee $O/ab_fuse_jasper10x5.txt
ls -la $O
