#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 600 python -m pytest tests/test_pathology_ops_gpu.py tests/test_gen_parity_gpu.py tests/test_native_planner_gpu.py -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/r2_tests11.log 2>&1
tail -15 gpurun_out/r2_tests11.log | cut -c1-220
for g in 0 2 4; do
  echo "finish group $g"; BFM_FINISH_GROUP=$g timeout 300 python tools/stage_bench.py 2>/dev/null | tee gpurun_out/r2_stage11_fg$g.json
done
for g in 0 4; do
  echo "finish group $g quick bench (sampler off)"; BFM_CLOCK_MS=0 BFM_FINISH_GROUP=$g timeout 300 python bench.py --steps 60 --warmup 5 --quick 2>/dev/null | cut -c1-120
done
# full-set capture of the non-warp kernels of one step (launch list of one step, then --set full on one launch each)
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_gen_upsample|k_gen_normalize|k_gen_band|k_gen_identity|k_gen_gmm_planes' -s 12 -c 12 -f -o gpurun_out/r2_full_rest python bench.py --steps 1 --warmup 3 --quick --lanes 1 > gpurun_out/r2_full_rest.log 2>&1
tail -2 gpurun_out/r2_full_rest.log
ncu -i gpurun_out/r2_full_rest.ncu-rep --page raw --csv > gpurun_out/r2_full_rest_raw.csv 2>/dev/null
ls -la gpurun_out/r2_full_rest*
