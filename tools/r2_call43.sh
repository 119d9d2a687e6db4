#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/r2_tests43.log 2>&1
tail -3 gpurun_out/r2_tests43.log | cut -c1-230
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench43.json 2> gpurun_out/r2_bench43.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench43.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'chain', d['chain']['frac_of_peak'])
print(d['chain']['host_wall_ms_per_step'], d['chain']['host_busy_ms_per_step'], d['cpu_baseline']['value'], d['gpu_launches'])
print({k: (v.get('total_ms'), v.get('frac_of_peak')) for k, v in d['configs'].items() if isinstance(v, dict)})
PY
tail -2 gpurun_out/r2_bench43.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
