#!/bin/bash
# final single-GPU pass of round 2: all GPU tests, the bench line, the launch list, one full capture of the warp kernel
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/r2_tests40.log 2>&1
tail -4 gpurun_out/r2_tests40.log | cut -c1-230
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench40.json 2> gpurun_out/r2_bench40.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench40.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'chain', d['chain']['frac_of_peak'])
print(d['chain']['stage_ms_per_step'], d['chain']['host_wall_ms_per_step'], d['chain']['host_busy_ms_per_step'])
print(d['cpu_baseline'])
print({k: (v.get('total_ms'), v.get('frac_of_peak')) for k, v in d['configs'].items() if isinstance(v, dict)})
PY
tail -2 gpurun_out/r2_bench40.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_v4.csv python bench.py --steps 2 --warmup 3 --quick --lanes 1 > gpurun_out/r2_launches_v4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_gen_warp_pk|k_gen_gmm_planes|k_gen_upsample|k_gen_normalize' -s 8 -c 4 -f -o gpurun_out/r2_full_v4 python bench.py --steps 1 --warmup 3 --quick --lanes 1 > gpurun_out/r2_full_v4.log 2>&1
ncu -i gpurun_out/r2_full_v4.ncu-rep --page raw --csv > gpurun_out/r2_full_v4_raw.csv 2>/dev/null
ls -la gpurun_out/r2_full_v4* gpurun_out/r2_launches_v4.csv
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-400
