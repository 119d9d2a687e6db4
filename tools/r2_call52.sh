#!/bin/bash
python -c 'import torch' >/dev/null 2>&1
timeout 600 python -m pytest tests/test_gen_parity_gpu.py -m gpu -q -x -p no:cacheprovider -k "slab" 2>&1 | tail -12 | cut -c1-300
