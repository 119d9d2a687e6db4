#!/bin/bash
python -c 'import torch' >/dev/null 2>&1
for i in 1 2; do timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['timed_region_samples'], d['clocks']['samples'], d['clocks']['under_load']['samples'], d['clocks']['reasons'])"; done
