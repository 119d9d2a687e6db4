#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 1500 python -m pytest tests/test_solvers_gpu.py tests/test_jitfields_compat.py tests/test_interpol_autograd_gpu.py tests/test_pipeline_gpu.py tests/test_shapeid_gpu.py -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/r2_tests4.log 2>&1
tail -4 gpurun_out/r2_tests4.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
cat gpurun_out/r2_bench4.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['chain']['stage_ms_per_step'], d['chain']['host_wall_ms_per_step'])"
tail -2 gpurun_out/r2_bench4.err
