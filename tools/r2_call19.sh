#!/bin/bash
# Round 2, call 19: whole GPU suite after the Jasper fp32 mode / host-path trims, smoke
O=gpurun_out/r2c19; mkdir -p $O
( time timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -s -k "fp32_faithful" ) 2>&1 | grep -E "tf32 mode|passed|failed" | cut -c1-300
( time timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider ) > $O/pytest_gpu.log 2>&1
tail -5 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
