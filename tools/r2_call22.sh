#!/bin/bash
# 8-GPU box: slab mode 512^3 on 8 / 4 ranks, then the bench line at N=8 (with the slab512 extra)
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
for n in 8 4; do
  SLAB_STEPS=10 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2971$n tools/slab_bench.py 2>gpurun_out/r2_slab_${n}gpu.err | tail -1 | tee gpurun_out/r2_slab_${n}gpu.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29720 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench22_8gpu.json 2> gpurun_out/r2_bench22_8gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench22_8gpu.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
print(json.dumps(d.get('configs'))[:800])
PY
tail -3 gpurun_out/r2_bench22_8gpu.err
