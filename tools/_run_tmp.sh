python -m pytest tests/test_gen_parity_gpu.py tests/test_native_planner_gpu.py -x -q 2>&1 | tail -6
python bench.py --no-cpu-baseline --quick
python tools/stage_bench.py 2>&1 | tail -1
