#!/bin/bash
# Profiling pass for one round (run under gpurun, ONE GPU):  tools/gpu_profile.sh <tag>
#   1. launch list of two bench steps (device time of every launch)       -> gpurun_out/launches_<tag>.csv
#   2. ncu --set full of one launch of each fused-chain kernel            -> gpurun_out/full_<tag>.ncu-rep
#   3. cProfile of the host side of generate_batch                        -> gpurun_out/host_<tag>.txt
# Numbers printed under ncu are never bench values.
tag=${1:-dev}
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/launches_${tag}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_gen_' \
    -s 60 -c 22 -f -o gpurun_out/full_${tag} python bench.py --steps 1 --warmup 3 --quick > gpurun_out/full_${tag}.log 2>&1
ncu -i gpurun_out/full_${tag}.ncu-rep --page raw --csv > gpurun_out/full_${tag}_raw.csv 2>/dev/null
timeout 300 python tools/host_profile.py > gpurun_out/host_${tag}.txt 2>&1
ls -la gpurun_out
