python -m pytest tests/test_interpol_gpu.py -x -q 2>&1 | tail -15
python tools/config_bench.py interpol 2>&1 | tail -3
