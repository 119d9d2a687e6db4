#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_native_planner_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -12 | cut -c1-250
for f in 1 0 1 0; do echo -n "fast=$f: "; BFM_FAST_SUBMIT=$f timeout 300 python bench.py --steps 60 --warmup 5 --quick 2>/dev/null | cut -c1-140; done
