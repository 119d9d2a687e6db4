python -m pytest tests/test_shapeid_gpu.py -x -q 2>&1 | tail -8
python tools/shapeid_profile.py 2>&1 | head -12
