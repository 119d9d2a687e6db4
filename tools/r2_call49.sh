#!/bin/bash
python -c 'import torch' >/dev/null 2>&1
for i in 1 2 3; do
  BFM_BENCH_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 5 --quick 2>&1 | grep -E "slow step|per-step" | cut -c1-420
done
