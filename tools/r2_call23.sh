#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 tools/symm_probe.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -8 | tee gpurun_out/r2_symm_probe.txt
SLAB_STEPS=20 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29732 tools/slab_bench.py 2>&1 | tail -1
