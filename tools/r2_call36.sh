#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 600 python tools/shapeid_profile.py > gpurun_out/r2_shapeid_profile.txt 2>&1
grep -v "^$" gpurun_out/r2_shapeid_profile.txt | head -70 | cut -c1-160
