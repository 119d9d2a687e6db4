#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 900 python -m pytest tests/test_shapeid_gpu.py tests/test_solvers_gpu.py tests/test_configs_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 | cut -c1-250
timeout 600 python tools/shapeid_profile.py 2>&1 | grep -E "rhs eval|dopri5:|GPU busy|k_advect" | cut -c1-150
timeout 600 python - <<'PY'
import sys, json
sys.path.insert(0, 'tools')
import config_bench as cb
print(json.dumps(cb.shapeid_cfg(192)))
PY
