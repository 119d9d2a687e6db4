#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 600 python -m pytest tests/test_gen_parity_gpu.py -m gpu -q -x -p no:cacheprovider -k "slab_mode_two_ranks" 2>&1 | tail -5 | cut -c1-300
SLAB_STEPS=10 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 tools/slab_bench.py 2>&1 | tail -2 | tee gpurun_out/r2_slab_2gpu.json
SLAB_STEPS=10 timeout 300 python tools/slab_bench.py 2>&1 | tail -1 | tee gpurun_out/r2_slab_1gpu.json
