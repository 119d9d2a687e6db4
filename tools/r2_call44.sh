#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29830 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench44_2gpu.json 2> gpurun_out/r2_bench44_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench44_2gpu.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['chain']['host_busy_ms_per_step'])
print(json.dumps(d.get('configs'))[:700])
PY
tail -2 gpurun_out/r2_bench44_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29831 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | cut -c1-200
