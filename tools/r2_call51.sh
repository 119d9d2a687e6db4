#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 600 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2 | cut -c1-200
for i in 1 2 3; do timeout 300 python bench.py --steps 20 --warmup 5 --quick 2>/dev/null | cut -c1-135; done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench51.json 2> gpurun_out/r2_bench51.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench51.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'chain', d['chain']['frac_of_peak'])
print(d['chain']['host_wall_ms_per_step'], d['chain']['host_busy_ms_per_step'], d['cpu_baseline']['value'], d['gpu_launches'])
print({k: (v.get('total_ms'), v.get('frac_of_peak')) for k, v in d['configs'].items() if isinstance(v, dict)})
PY
