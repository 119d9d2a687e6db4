#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_spline_filter_tile' -s 3 -c 3 -f -o gpurun_out/r2_full_prefilter python tools/prefilter_probe.py > gpurun_out/r2_full_prefilter.log 2>&1
tail -2 gpurun_out/r2_full_prefilter.log
ncu -i gpurun_out/r2_full_prefilter.ncu-rep --page raw --csv > gpurun_out/r2_full_prefilter_raw.csv 2>/dev/null
ls -la gpurun_out/r2_full_prefilter*
