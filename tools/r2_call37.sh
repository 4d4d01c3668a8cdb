#!/bin/bash
# Round 2, call 37: BatchNorm statistics in the forward GEMM epilogue (default) vs a separate pass over z, with the CTA-pair kernels
bash tools/ab.sh W2L_EPILOGUE_STATS 1 0
bash tools/ab.sh W2L_EPILOGUE_STATS 1 0 --model jasper10x5 --steps 10
