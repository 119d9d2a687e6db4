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
    # phase timers without profiler overhead
    from brainfm_b200.Generator import datasets as D, utils as U
    acc = {}

    def wrap(owner, name):
        fn = getattr(owner, name)

        def w(*a, **k):
            t = time.perf_counter_ns()
            try:
                return fn(*a, **k)
            finally:
                acc[name] = acc.get(name, 0) + time.perf_counter_ns() - t
        setattr(owner, name, w)
        return fn

    saved = []
    for owner, names in ((D.BaseGen, ["_prologue_host", "read_input", "get_setup_params", "generate_deformation",
                                      "random_affine_transform", "random_nonlinear_transform", "_plan_synth",
                                      "get_contrast", "_build_descs", "_run_chain", "_targets", "_finish_item",
                                      "_fused_image_targets", "_job"]),
                         (U.DeformPlan, ["__init__"]), (U, ["make_affine_matrix"])):
        for nm in names:
            saved.append((owner, nm, wrap(owner, nm)))
    D.make_affine_matrix = U.make_affine_matrix
    t0 = time.perf_counter()
    for _ in range(n):
        ds.generate_batch(idxs)
    tot = time.perf_counter() - t0
    torch.cuda.synchronize()
    print("phase timers (us per sample, inclusive): total %.1f" % (1e6 * tot / n / bench.BATCH))
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
        print("  %-28s %7.1f" % (k, v / 1e3 / n / bench.BATCH))
    for owner, nm, fn in saved:
        setattr(owner, nm, fn)
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
