#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/r2_tests53.log 2>&1
tail -3 gpurun_out/r2_tests53.log | cut -c1-230
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
