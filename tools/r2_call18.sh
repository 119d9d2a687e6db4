#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 900 python -m pytest tests/test_gen_parity_gpu.py -m gpu -q -x -p no:cacheprovider -k "slab" 2>&1 | tail -5 | cut -c1-300
SLAB_STEPS=10 timeout 600 python tools/slab_bench.py 2>&1 | tail -1
timeout 600 python tools/slab_hostprof.py > gpurun_out/r2_slab_hostprof3.txt 2>&1
head -c 3800 gpurun_out/r2_slab_hostprof3.txt
