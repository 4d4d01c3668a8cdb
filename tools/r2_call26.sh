#!/bin/bash
# Round 2, call 26: compute-sanitizer over the kernels written this round (CTA-pair GEMMs, tf32 path, fused reduction epilogue, padded
# widths, fp32 depthwise), caching allocator off so that every tensor is its own allocation
O=gpurun_out/r2c26; mkdir -p $O
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
timeout 900 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -q -m gpu -p no:cacheprovider \
   -k "(conv or tf32 or fused_bn_reduce or fp32_faithful or odd_widths or bn_act or bn_passes) and not full_size and not slab and not config1" > $O/sanitizer_memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $O/sanitizer_memcheck.log | tail -8
timeout 600 compute-sanitizer --tool initcheck --print-limit 30 python -m pytest tests/test_gpu_models.py tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider \
   -k "fp32_faithful or odd_widths or w2l_golden or fused_bn_reduce or tf32" > $O/sanitizer_initcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed|Uninitialized" $O/sanitizer_initcheck.log | tail -8
