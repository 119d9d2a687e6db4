#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_gen_upsample' -s 3 -c 1 -f -o gpurun_out/r2_full_ups python bench.py --steps 1 --warmup 3 --quick --lanes 1 > gpurun_out/r2_full_ups.log 2>&1
tail -2 gpurun_out/r2_full_ups.log
ncu -i gpurun_out/r2_full_ups.ncu-rep --page raw --csv > gpurun_out/r2_full_ups_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_full_ups.ncu-rep --page source --csv > gpurun_out/r2_full_ups_src.csv 2>/dev/null
ls -la gpurun_out/r2_full_ups*
