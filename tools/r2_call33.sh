#!/bin/bash
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
timeout 900 python -m pytest tests/test_interpol_gpu.py tests/test_configs_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 | cut -c1-250
timeout 600 python - <<'PY'
import sys, json
sys.path.insert(0, 'tools')
import config_bench as cb
r = cb.interpol_cfg(256)
print(json.dumps(r))
PY
