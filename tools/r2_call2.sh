#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
for v in pk5 pk6; do
  BFM_LIB=$PWD/brainfm_b200/libbfm_$v.so timeout 300 python tools/stage_bench.py 2>/dev/null | tee gpurun_out/r2_stage_$v.json
done
for g in 8 4 2 1; do
  BFM_GEN_GROUP=$g timeout 300 python tools/simple_bench.py 2>/dev/null | tee -a gpurun_out/r2_simple.jsonl
done
for g in 8 2 1; do
  STREAMS=2 BFM_GEN_GROUP=$g timeout 300 python tools/simple_bench.py 2>/dev/null | tee -a gpurun_out/r2_simple.jsonl
done
STREAMS=3 BFM_GEN_GROUP=1 timeout 300 python tools/simple_bench.py 2>/dev/null | tee -a gpurun_out/r2_simple.jsonl
