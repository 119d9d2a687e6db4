python -m pytest tests/test_gen_parity_gpu.py tests/test_surface.py -x -q 2>&1 | tail -15
