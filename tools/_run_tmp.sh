python -m pytest tests/test_interpol_autograd_gpu.py tests/test_interpol_gpu.py -x -q 2>&1 | tail -15
