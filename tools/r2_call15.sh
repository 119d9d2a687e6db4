#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
for t in 1 2 4 8; do
  echo "gmm tiles $t"; BFM_GMM_TILES=$t timeout 300 python tools/stage_bench.py 2>/dev/null
done
