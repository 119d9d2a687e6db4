#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_advect_rhs_p' -s 20 -c 2 -f -o gpurun_out/r2_full_advect python tools/shapeid_profile.py > gpurun_out/r2_full_advect.log 2>&1
tail -2 gpurun_out/r2_full_advect.log
ncu -i gpurun_out/r2_full_advect.ncu-rep --page raw --csv > gpurun_out/r2_full_advect_raw.csv 2>/dev/null
ls -la gpurun_out/r2_full_advect*
