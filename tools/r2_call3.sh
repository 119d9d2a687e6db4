#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/r2_tests3.log 2>&1
tail -4 gpurun_out/r2_tests3.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
cat gpurun_out/r2_bench3.json; tail -3 gpurun_out/r2_bench3.err
timeout 200 python tools/hostlink_probe.py > gpurun_out/r2_hostlink_1gpu.json 2>gpurun_out/r2_hostlink_1gpu.err; cat gpurun_out/r2_hostlink_1gpu.json
