#!/bin/bash
# Round 2, call 12: tf32 GEMMs (fp32-faithful mode) on hardware
O=gpurun_out/r2c12; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "tf32 or conv" ) > $O/pytest_tf32.log 2>&1
tail -30 $O/pytest_tf32.log
