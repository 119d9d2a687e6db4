#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 900 python -m pytest tests/test_gen_parity_gpu.py tests/test_native_planner_gpu.py tests/test_pipeline_gpu.py tests/test_shapeid_gpu.py tests/test_solvers_gpu.py tests/test_noise_gpu.py -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/r2_tests5.log 2>&1
tail -25 gpurun_out/r2_tests5.log | cut -c1-220
timeout 300 python -m pytest tests/test_interpol_autograd_gpu.py -m gpu -q -k "adjoints or gradcheck_grad" --maxfail=10 -p no:cacheprovider > gpurun_out/r2_tests5b.log 2>&1
tail -8 gpurun_out/r2_tests5b.log | cut -c1-220
for v in tile2 tile3 pk; do
  case $v in
    tile2) env="";;
    tile3) env="BFM_LIB=$PWD/brainfm_b200/libbfm_t3.so";;
    pk) env="BFM_WARP_TILE=0";;
  esac
  env $env timeout 300 python tools/stage_bench.py 2>/dev/null | tee gpurun_out/r2_stage5_$v.json
done
BFM_LIB=$PWD/brainfm_b200/libbfm_t3.so BFM_BRICK_KB=40 timeout 300 python tools/stage_bench.py 2>/dev/null | tee gpurun_out/r2_stage5_tile3_40.json
BFM_BRICK_KB=64 timeout 300 python tools/stage_bench.py 2>/dev/null | tee gpurun_out/r2_stage5_tile2_64.json
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
cat gpurun_out/r2_bench5.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['chain']['stage_ms_per_step'], d['chain']['host_wall_ms_per_step'])"
tail -2 gpurun_out/r2_bench5.err
