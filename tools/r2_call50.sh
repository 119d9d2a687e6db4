#!/bin/bash
python -c 'import torch' >/dev/null 2>&1
BFM_BENCH_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 5 --quick 2>&1 | grep -E "^step|slow step 3" | cut -c1-300
