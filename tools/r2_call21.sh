#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
for i in 1 2 3; do
  timeout 600 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('full', round(d['value']), d['ms_per_step'], d['clocks'].get('sampler'), d['clocks']['samples'])"
  timeout 300 python bench.py --steps 30 --warmup 5 --quick 2>/dev/null | cut -c1-60
done
BFM_CLOCK_MS=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('full nosampler', round(d['value']), d['ms_per_step'])"
