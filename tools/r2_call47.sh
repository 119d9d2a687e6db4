#!/bin/bash
python -c 'import torch' >/dev/null 2>&1
for f in 1 0; do
  echo "fast=$f"; BFM_BENCH_TRACE=1 BFM_FAST_SUBMIT=$f timeout 300 python bench.py --steps 40 --warmup 5 --quick 2>&1 | grep -E "per-step|quick" | cut -c1-700
done
