"""Host-side profile of DevicePipeline.submit (development tool)."""
import cProfile
import io
import json
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from brainfm_b200.pipeline import DevicePipeline


def loop(pipe, idxs, steps, lanes):
    tickets = []
    for _ in range(steps):
        tickets.append(pipe.submit(idxs))
        if len(tickets) > lanes:
            tickets.pop(0).wait()
    for t in tickets:
        t.wait()


def main():
    dev = torch.device("cuda", 0)
    ds = bench.build_dataset(bench.make_inputs(2 * bench.BATCH), dev)
    np.random.seed(1000)
    idxs = list(range(bench.BATCH))
    out = {}
    for lanes in (1, 2, 3):
        pipe = DevicePipeline(ds, depth=lanes)
        loop(pipe, idxs, 10, lanes)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        loop(pipe, idxs, 100, lanes)
        t1 = time.perf_counter()          # host time to ENQUEUE 100 steps (may include waiting for arena slots)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        out["lanes%d" % lanes] = {"enqueue_ms_per_step": 10 * (t1 - t0), "total_ms_per_step": 10 * (t2 - t0)}
    # plain loop for comparison
    for _ in range(10):
        ds.generate_batch(idxs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(100):
        ds.generate_batch(idxs)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    out["plain"] = {"enqueue_ms_per_step": 10 * (t1 - t0), "total_ms_per_step": 10 * (t2 - t0)}
    print(json.dumps(out))
    pipe = DevicePipeline(ds, depth=2)
    loop(pipe, idxs, 10, 2)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    loop(pipe, idxs, 200, 2)
    pr.disable()
    torch.cuda.synchronize()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
    print(s.getvalue()[:6000])


if __name__ == "__main__":
    main()
