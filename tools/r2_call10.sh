#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
for ms in 0 20 200 20 0; do
  echo "sampler $ms"; BFM_CLOCK_MS=$ms timeout 300 python bench.py --steps 30 --warmup 5 --quick 2>/dev/null | cut -c1-400
done
for ms in 0 20; do
  echo "sampler $ms, 100 steps"; BFM_CLOCK_MS=$ms timeout 300 python bench.py --steps 100 --warmup 5 --quick 2>/dev/null | cut -c1-400
done
