#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
for v in "" _u16 _u48 _u32m6 _u32m10; do
  echo "variant libbfm$v.so"; BFM_LIB=$PWD/brainfm_b200/libbfm$v.so timeout 300 python tools/stage_bench.py 2>/dev/null
done
timeout 300 python -m pytest tests/test_gen_parity_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
