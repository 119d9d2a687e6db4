"""cProfile of the host side of BaseGen.generate_batch (bench configuration).  Development tool."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench


def main():
    dev = torch.device("cuda", 0)
    ds = bench.build_dataset(bench.make_inputs(bench.BATCH), dev)
    np.random.seed(1000)
    torch.manual_seed(1000)
    idxs = list(range(bench.BATCH))
    for _ in range(5):
        ds.generate_batch(idxs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 30
    for _ in range(n):
        ds.generate_batch(idxs)
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    print("host ms/batch (no profiler): %.3f   wall incl. drain: %.3f" % (1e3 * t_host / n,
                                                                        1e3 * (time.perf_counter() - t0) / n))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        ds.generate_batch(idxs)
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr, stream=sys.stdout)
    st.sort_stats("cumulative").print_stats(45)
    st.sort_stats("tottime").print_stats(35)


if __name__ == "__main__":
    main()
