#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 1200 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --durations=5 > gpurun_out/r2_tests9.log 2>&1
tail -25 gpurun_out/r2_tests9.log | cut -c1-230
timeout 300 python tools/pipe_profile.py > gpurun_out/r2_pipe_profile9.txt 2>&1
head -c 1500 gpurun_out/r2_pipe_profile9.txt
timeout 600 python bench.py --steps 30 --warmup 5 --no-extras > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench9.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'chain', d['chain']['frac_of_peak'])
print(d['chain']['stage_ms_per_step'], d['chain']['host_wall_ms_per_step'])
PY
tail -3 gpurun_out/r2_bench9.err
