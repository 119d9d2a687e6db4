#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 600 python -m pytest tests/test_gen_parity_gpu.py tests/test_native_planner_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4 | cut -c1-250
for b in 0 1; do echo "bulk $b"; BFM_UPSAMPLE_BULK=$b timeout 300 python tools/stage_bench.py 2>/dev/null; done
for b in 0 1; do echo -n "bulk $b: "; BFM_UPSAMPLE_BULK=$b timeout 300 python bench.py --steps 60 --warmup 5 --quick 2>/dev/null | cut -c1-60; done
