nvidia-smi topo -m 2>&1 | head -6
lscpu | grep -i "numa\|^CPU(s)"
for v in 0 1; do
  if [ $v = 0 ]; then export BFM_NUMA_BIND=1; else unset BFM_NUMA_BIND; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$v bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline 2> gpurun_out/b2_$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bind=%s' % ('no' if $v else 'yes'), round(d['value']), round(d['e2e']['value']))"
done
