#!/bin/bash
mkdir -p gpurun_out
tools/micro/texgather > gpurun_out/r2_texgather.txt 2>&1
cat gpurun_out/r2_texgather.txt
python -c 'import torch' >/dev/null 2>&1
timeout 600 python -m pytest tests/test_gen_parity_gpu.py -m gpu -q -x -p no:cacheprovider -k "slab" 2>&1 | tail -15 | cut -c1-250
