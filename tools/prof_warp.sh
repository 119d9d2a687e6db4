mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_gen_warp' -s 3 -c 1 -f -o gpurun_out/warp_v6 python bench.py --steps 1 --warmup 3 --quick > gpurun_out/warp_v6.log 2>&1
tail -3 gpurun_out/warp_v6.log
ls -la gpurun_out/*.ncu-rep
