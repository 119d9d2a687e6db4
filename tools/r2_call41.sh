#!/bin/bash
python -c 'import torch' >/dev/null 2>&1
echo "A: bench --no-cpu-baseline"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(d['configs']['shapeid192']['ms'])"
echo "B: interpol_cfg then shapeid_cfg in one process"
timeout 600 python - <<'PY'
import sys
sys.path.insert(0, 'tools')
import config_bench as cb
cb.interpol_cfg(256)
print(cb.shapeid_cfg(192)['ms'])
print(cb.shapeid_cfg(192)['ms'])
PY
echo "C: oracle CPU baseline first (torch CPU threads), then shapeid"
timeout 600 python - <<'PY'
import sys, os
sys.path.insert(0, 'tools')
import bench, torch
subs = bench.make_inputs(2)
gens = bench.oracle_generator(subs[:2], os.cpu_count())
bench.time_oracle(gens, 2, seed=5)
import config_bench as cb
print(cb.shapeid_cfg(192)['ms'], torch.get_num_threads())
PY
