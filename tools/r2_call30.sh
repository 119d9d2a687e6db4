#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 1200 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/r2_tests30.log 2>&1
tail -6 gpurun_out/r2_tests30.log | cut -c1-230
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench30.json 2> gpurun_out/r2_bench30.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench30.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'chain', d['chain']['frac_of_peak'], d['chain']['nc_over_n'], d['chain']['ns_over_n'])
print(d['chain']['stage_ms_per_step'], d['chain']['host_wall_ms_per_step'])
print(d['cpu_baseline'])
PY
tail -3 gpurun_out/r2_bench30.err
timeout 300 python tools/stage_bench.py 2>/dev/null
