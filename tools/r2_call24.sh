#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
SLAB_PROFILE=1 SLAB_STEPS=10 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29741 tools/slab_bench.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -20 | tee gpurun_out/r2_slab_profile_4gpu.txt
