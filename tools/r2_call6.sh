#!/bin/bash
# 8-GPU box: host-link ceiling with / without NUMA binding, then the bench both ways
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
lscpu | head -30 > gpurun_out/r2_lscpu.txt 2>&1
numactl -H >> gpurun_out/r2_lscpu.txt 2>&1
for n in 8 4 2; do
for b in 0 1; do
  BFM_NUMA_BIND=$b timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 tools/hostlink_probe.py 2>gpurun_out/r2_hostlink_${n}gpu_bind$b.err | tee gpurun_out/r2_hostlink_${n}gpu_bind$b.json
done
done
for b in 0 1; do
  BFM_NUMA_BIND=$b timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench8_bind$b.json 2> gpurun_out/r2_bench8_bind$b.err
  python -c "import json; d=json.load(open('gpurun_out/r2_bench8_bind$b.json')); print('bind$b', d['value'], d['ms_per_step'], d['e2e'])"
done
