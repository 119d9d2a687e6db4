"""Kernel-time breakdown of one 512^3 volume in slab mode on one GPU (torch profiler, CUDA activities) and the host
time of generate_slab.  Development tool."""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from brainfm_b200 import io as bio
from brainfm_b200.Generator import BaseGen
from brainfm_b200.Generator.slab import generate_slab
from tests import _inputs as ti


def main():
    size = int(os.environ.get("SLAB_SIZE", "512"))
    dev = torch.device("cuda", 0)
    half = ti.brain_like_labels((size // 2,) * 3, seed=7)
    lab = np.repeat(np.repeat(np.repeat(half, 2, 0), 2, 1), 2, 2)
    root = tempfile.mkdtemp(prefix="bfm_slab_")
    stem = os.path.join(root, "HCP.sub00.")
    bio.register_volume(stem + "T1w.nii", np.zeros((2, 2, 2), dtype=np.float32))
    bio.register_volume(stem + "generation_labels.nii", lab)
    with open(os.path.join(root, "train.txt"), "w") as f:
        f.write(stem + "T1w.nii\n")
    cfg = ti.default_cfg((size,) * 3)
    for k in vars(cfg.task):
        setattr(cfg.task, k, False)
    cfg.split_root = root
    ds = BaseGen(cfg, dev, planner='python')
    ds.write_bflog = True
    np.random.seed(4321)
    torch.manual_seed(4321)
    for _ in range(2):
        generate_slab(ds, 0, 0, 1)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    n = 4
    t0 = time.perf_counter()
    host = 0.0
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            h0 = time.perf_counter()
            generate_slab(ds, 0, 0, 1)
            host += time.perf_counter() - h0
        torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    print("wall %.2f ms/volume, host inside generate_slab %.2f ms/volume" % (1e3 * wall / n, 1e3 * host / n))
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:16]
    tot = sum(e.device_time_total for e in prof.key_averages())
    print("device time %.2f ms/volume" % (tot / n / 1e3))
    for e in rows:
        print("%8.3f ms  x%-3d %s" % (e.device_time_total / n / 1e3, e.count // n, e.key[:90]))


if __name__ == "__main__":
    main()
