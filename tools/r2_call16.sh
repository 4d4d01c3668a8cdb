#!/bin/bash
# Round 2, call 16: CTC schedules at the spill-bound corner: parallel (alpha || beta, both spilled) vs serial (alpha spilled, beta fused with the gradient)
O=gpurun_out/r2c16; mkdir -p $O
for v in 0 8 0 8; do echo "## W2L_CTC_DBG=$v"; W2L_CTC_DBG=$v timeout 300 python tools/sweep_ctc_decode.py --quick; done 2>&1 | tee $O/ctc_serial_ab.md
