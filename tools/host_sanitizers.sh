#!/bin/bash
# memcheck / initcheck / alignment check of the kernels WITHOUT a GPU: the `-m gpu` suite on the emulated GPU (tests/_fake_cuda.py)
# with the kernels' host builds instrumented by AddressSanitizer, then by UndefinedBehaviorSanitizer.
#   ASan : tensors come from ASan's allocator with red zones, so a kernel access past a buffer aborts with file:line of the .cu text
#   UBSan: CUDA's vector types keep their alignment on the host, so a misaligned 8 / 16-byte access (a fault on the GPU) is reported,
#          as are signed overflow / bad shifts in index arithmetic
#   schedule: the fibers of a block resumed in reverse / random order every round -- results must not change (racecheck in spirit)
# (uninitialised reads are covered by the poisoned allocations every emulated run uses.)  About 5 minutes each on 8 cores.
set -e
cd "$(dirname "$0")/.."
echo "== AddressSanitizer"
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 W2L_EMU_ASAN=1 \
  python -m pytest tests -m gpu --emulate-gpu -q -p no:cacheprovider -x 2>&1 | tail -3
echo "== UndefinedBehaviorSanitizer"
W2L_EMU_UBSAN=1 python -m pytest tests -m gpu --emulate-gpu -q -p no:cacheprovider -x 2>&1 | tail -3
echo "== thread schedule: reverse, random"
for m in reverse random:1 random:2; do
  W2L_EMU_SCHEDULE=$m python -m pytest tests -m gpu --emulate-gpu -q -p no:cacheprovider -x 2>&1 | tail -1
done
