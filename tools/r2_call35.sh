#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 300 python tools/stage_bench.py 2>/dev/null
timeout 900 python -m pytest tests/test_gen_parity_gpu.py tests/test_native_planner_gpu.py tests/test_noise_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 | cut -c1-250
for i in 1 2; do timeout 300 python bench.py --steps 60 --warmup 5 --quick 2>/dev/null | cut -c1-150; done
