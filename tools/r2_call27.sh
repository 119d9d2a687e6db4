#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
for l in 1 2; do
SLAB_LANES=$l SLAB_STEPS=20 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2977$l tools/slab_bench.py 2>&1 | tail -1 | cut -c90-420
SLAB_LANES=$l SLAB_STEPS=20 timeout 300 python tools/slab_bench.py 2>&1 | tail -1 | cut -c90-420
done
