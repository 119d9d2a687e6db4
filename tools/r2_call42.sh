#!/bin/bash
# 8-GPU box, final: slab mode 512^3 on 8 / 4 ranks (two volumes in flight, same 20 volumes), bench line at N=8
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
for n in 8 4; do
  SLAB_STEPS=20 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2981$n tools/slab_bench.py 2>gpurun_out/r2_slab_v2_${n}gpu.err | tail -1 | tee gpurun_out/r2_slab_v2_${n}gpu.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29820 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench42_8gpu.json 2> gpurun_out/r2_bench42_8gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench42_8gpu.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
print(json.dumps(d.get('configs'))[:800])
PY
tail -2 gpurun_out/r2_bench42_8gpu.err
