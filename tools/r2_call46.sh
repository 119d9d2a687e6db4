#!/bin/bash
python -c 'import torch' >/dev/null 2>&1
for f in 1 0 1 0; do
  echo -n "full fast=$f: "; BFM_FAST_SUBMIT=$f timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); c=d['chain']; print(round(d['value']), round(d['ms_per_step'],3), 'busy', round(c['host_busy_ms_per_step'],3), 'wall', round(c['host_wall_ms_per_step'],3), d['gpu_launches'])"
done
for f in 1 0; do echo -n "quick fast=$f: "; BFM_FAST_SUBMIT=$f timeout 300 python bench.py --steps 20 --warmup 5 --quick 2>/dev/null | cut -c1-140; done
nproc; cat /proc/loadavg
