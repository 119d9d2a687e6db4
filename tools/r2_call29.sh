#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 300 python tools/stage_bench.py 2>/dev/null
echo "unroll 8"; BFM_LIB=$PWD/brainfm_b200/libbfm_bu8.so timeout 300 python tools/stage_bench.py 2>/dev/null
for b in 8 24 32; do echo "band blocks/SM $b"; BFM_BAND_BLOCKS_PER_SM=$b timeout 300 python tools/stage_bench.py 2>/dev/null; done
timeout 600 python -m pytest tests/test_gen_parity_gpu.py -m gpu -q -x -p no:cacheprovider -k "bulk" 2>&1 | tail -2
