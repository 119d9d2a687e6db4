#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 600 python tools/slab_hostprof.py > gpurun_out/r2_slab_hostprof.txt 2>&1
head -c 5000 gpurun_out/r2_slab_hostprof.txt
