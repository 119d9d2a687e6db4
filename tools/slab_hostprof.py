"""cProfile of generate_slab's host side (one 512^3 volume, one GPU).  Development tool."""
import cProfile
import io
import os
import pstats
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
    world = int(os.environ.get("SLAB_FAKE_WORLD", "1"))       # pretend to be rank 0 of `world` (no exchange partner)
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
    ds = BaseGen(cfg, dev, planner=os.environ.get('SLAB_PLANNER', 'auto'))
    ds.write_bflog = True
    np.random.seed(4321)
    torch.manual_seed(4321)
    for _ in range(3):
        generate_slab(ds, 0, 0, 1)
    torch.cuda.synchronize()
    n = 20
    t0 = time.perf_counter()
    for _ in range(n):
        generate_slab(ds, 0, 0, 1)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("enqueue %.3f ms/volume, total %.3f ms/volume" % (1e3 * (t1 - t0) / n, 1e3 * (t2 - t0) / n))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        generate_slab(ds, 0, 0, 1)
    pr.disable()
    torch.cuda.synchronize()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35)
    print(s.getvalue()[:6000])


if __name__ == "__main__":
    main()
