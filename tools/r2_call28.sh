#!/bin/bash
# Round 2, call 28: does replaying the W2L-20 headline step from a CUDA graph buy anything (it is not launch-bound)?
timeout 600 python tools/graph_vs_eager.py 20 64 15 20 2>&1 | grep -v Warning | tail -4
timeout 600 python tools/graph_vs_eager.py 5 16 5 50 2>&1 | grep -v Warning | tail -4
