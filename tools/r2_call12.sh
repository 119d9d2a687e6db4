#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 600 python -m pytest tests/test_gen_parity_gpu.py tests/test_native_planner_gpu.py tests/test_pipeline_gpu.py tests/test_noise_gpu.py -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/r2_tests12.log 2>&1
tail -8 gpurun_out/r2_tests12.log | cut -c1-220
timeout 300 python tools/stage_bench.py 2>/dev/null | tee gpurun_out/r2_stage12.json
for i in 1 2; do BFM_CLOCK_MS=0 timeout 300 python bench.py --steps 60 --warmup 5 --quick 2>/dev/null | cut -c1-100; done
