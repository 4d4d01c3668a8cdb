#!/bin/bash
# Round 2, call 13: fp32-faithful mode (tf32 GEMMs, fp32 BatchNorm passes) on hardware + whole suite
O=gpurun_out/r2c13; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -q -m gpu -p no:cacheprovider -k "tf32 or fp32_faithful" -s ) > $O/pytest_tf32.log 2>&1
tail -12 $O/pytest_tf32.log | cut -c1-300
( time timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider ) > $O/pytest_gpu.log 2>&1
tail -6 $O/pytest_gpu.log
