#!/bin/bash
show='import sys, json
l = json.loads(sys.stdin.readline()); print(sys.argv[1], "main", round(l["ms_per_step"],2), "default", l.get("default_config"))'
timeout 600 python bench.py --steps 20 --warmup 4 --skip-cpu --skip-legs 2>/dev/null | tail -1 | python -c "$show" full
W2L_BENCH_SKIP_ISO=1 timeout 600 python bench.py --steps 20 --warmup 4 --skip-cpu --skip-legs 2>/dev/null | tail -1 | python -c "$show" skip_iso
timeout 600 python bench.py --steps 5 --warmup 3 --mid-layers 3 --skip-cpu --skip-legs 2>/dev/null | tail -1 | python -c "$show" mid3
