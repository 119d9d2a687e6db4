#!/bin/bash
python -c 'import torch' >/dev/null 2>&1
timeout 900 python -m pytest tests/test_dopri5_interp_gpu.py tests/test_shapeid_gpu.py tests/test_solvers_gpu.py tests/test_configs_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5 | cut -c1-250
