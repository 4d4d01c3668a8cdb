#!/bin/bash
# Round 2, call 43: graph step guards (replaced optimizer state) -- the graph test file
( timeout 900 python -m pytest tests/test_gpu_zz_graph.py -q -m gpu -p no:cacheprovider ) 2>&1 | grep -E "passed|failed|FAILED|Error" | tail -6 | cut -c1-300
