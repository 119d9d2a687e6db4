python bench.py > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err
cat gpurun_out/bench_v6.json
python tools/stage_bench.py 2>&1 | tail -1 > gpurun_out/stage_v6.json; cat gpurun_out/stage_v6.json
python tools/config_bench.py > gpurun_out/configs_v6.json 2>&1; cat gpurun_out/configs_v6.json
