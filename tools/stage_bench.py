"""Times each stage of the fused chain with the GPU saturated: one batch is planned once, then every stage is
launched REPS times back to back between two CUDA events.  Development tool (not a bench value)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from brainfm_b200 import _lib

REPS = int(os.environ.get("REPS", "20"))


def main():
    dev = torch.device("cuda", 0)
    ds = bench.build_dataset(bench.make_inputs(bench.BATCH), dev)
    np.random.seed(1000)
    torch.manual_seed(1000)
    for _ in range(3):
        ds.generate_batch(list(range(bench.BATCH)))
    torch.cuda.synchronize()
    descs, d_dev, B = ds._last_descs
    L = _lib.lib()
    h = C.addressof(descs)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {}
    for name in ("bbox", "gmm", "warp", "resample", "finish"):
        fn = getattr(L, "bfm_gen_" + name)
        fn(h, d_dev, B, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(REPS):
            _lib.check(fn(h, d_dev, B, st))
        e1.record()
        torch.cuda.synchronize()
        out[name] = round(e0.elapsed_time(e1) / REPS * 1e3 / B, 2)     # us per sample
    out["total_us_per_sample"] = round(sum(out.values()), 2)
    out["lib"] = os.path.basename(_lib.LIB_PATH)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
