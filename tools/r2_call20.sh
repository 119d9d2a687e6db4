#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
for mode in none nvml smi nvml smi none; do
  case $mode in
    none) e="BFM_CLOCK_MS=0";; nvml) e="BFM_CLOCK_MS=20";; smi) e="BFM_CLOCK_SMI=1";;
  esac
  echo -n "sampler $mode: "; env $e timeout 300 python bench.py --steps 30 --warmup 5 --quick 2>/dev/null | cut -c1-60
done
timeout 1200 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/r2_tests20.log 2>&1
tail -6 gpurun_out/r2_tests20.log | cut -c1-230
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench20.json 2> gpurun_out/r2_bench20.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench20.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'chain', d['chain']['frac_of_peak'])
print(d['chain']['stage_ms_per_step'], d['chain']['host_wall_ms_per_step'])
print(d['clocks'])
print(json.dumps(d.get('configs'))[:1200])
PY
tail -3 gpurun_out/r2_bench20.err
