#!/bin/bash
# round 2, GPU call 1: full GPU test suite + stage timings of the pair-mode warp kernel variants
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/r2_tests1.log 2>&1
tail -5 gpurun_out/r2_tests1.log
for v in default nopair pk2 pk4; do
  case $v in
    default) env="";;
    nopair) env="BFM_PAIR_MODE=0";;
    pk2) env="BFM_LIB=$PWD/brainfm_b200/libbfm_pk2.so";;
    pk4) env="BFM_LIB=$PWD/brainfm_b200/libbfm_pk4.so";;
  esac
  env $env timeout 300 python tools/stage_bench.py > gpurun_out/r2_stage_$v.json 2> gpurun_out/r2_stage_$v.err
  echo "$v: $(cat gpurun_out/r2_stage_$v.json)"
done
timeout 300 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2_quick1.json 2> gpurun_out/r2_quick1.err
cat gpurun_out/r2_quick1.json
