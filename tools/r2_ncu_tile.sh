#!/bin/bash
tag=${1:-tile}
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_gen_warp_tile' -s 3 -c 1 -f -o gpurun_out/r2_full_${tag} python bench.py --steps 1 --warmup 3 --quick > gpurun_out/r2_full_${tag}.log 2>&1
tail -2 gpurun_out/r2_full_${tag}.log
ncu -i gpurun_out/r2_full_${tag}.ncu-rep --page raw --csv > gpurun_out/r2_full_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_full_${tag}.ncu-rep --page source --csv > gpurun_out/r2_full_${tag}_src.csv 2>/dev/null
ls -la gpurun_out/r2_full_${tag}*
