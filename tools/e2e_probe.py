"""Probe of the end-to-end arm: raw PCIe copy rates from pinned memory (each direction alone and both at once),
then host-side and device-side timing of HostPipeline.submit.  Development tool."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench


def copy_rates(dev):
    n = 160 ** 3
    h_in = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(8)]
    h_out = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(8)]
    d_in = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(8)]
    d_out = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(8)]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(up, down, reps=5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s1):
                    for a, b in zip(d_in, h_in):
                        a.copy_(b, non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    for a, b in zip(h_out, d_out):
                        a.copy_(b, non_blocking=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    run(True, True, 2)
    mb = 8 * n * 4 / 1e6
    a, b, c = run(True, False), run(False, True), run(True, True)
    print("H2D %.0f MB: %.2f ms (%.1f GB/s)   D2H: %.2f ms (%.1f GB/s)   both at once: %.2f ms" %
          (mb, a * 1e3, mb / a / 1e3, b * 1e3, mb / b / 1e3, c * 1e3))


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    copy_rates(dev)
    from brainfm_b200.pipeline import HostPipeline
    B = bench.BATCH
    subs = bench.make_inputs(2 * B)
    ds = bench.build_dataset(subs, dev)
    np.random.seed(1000)
    torch.manual_seed(1000)
    sets = [list(range(0, B)), list(range(B, 2 * B))]
    if "--per-volume" in sys.argv:      # one DMA per volume (16 per step) instead of one per kind (2 per step)
        host_lab = [torch.from_numpy(s["Gen"].astype(np.uint8)).pin_memory() for s in subs]
        host_t1 = [torch.from_numpy(s["T1"]).pin_memory() for s in subs]
        uploads = [[u for s in st for u in ((ds.names[0][s][:-7] + "generation_labels.nii", "gen", host_lab[s]),
                                            (ds.names[0][s], "f32", host_t1[s]))] for st in sets]
    else:
        host_lab = [torch.from_numpy(np.stack([subs[s]["Gen"].astype(np.uint8) for s in st])).pin_memory()
                    for st in sets]
        host_t1 = [torch.from_numpy(np.stack([subs[s]["T1"] for s in st])).pin_memory() for st in sets]
        uploads = [[([ds.names[0][s][:-7] + "generation_labels.nii" for s in st], "gen", host_lab[q]),
                    ([ds.names[0][s] for s in st], "f32", host_t1[q])] for q, st in enumerate(sets)]
    pipe = HostPipeline(ds, depth=3)
    if "--trace" in sys.argv:
        from torch.profiler import profile, ProfilerActivity
        tickets = []
        for k in range(4):
            tickets.append(pipe.submit(sets[k % 2], uploads[k % 2]))
        for t in tickets:
            t.wait()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            tickets = []
            for k in range(6):
                tickets.append(pipe.submit(sets[k % 2], uploads[k % 2]))
                if len(tickets) > 2:
                    tickets.pop(0).wait()
            for t in tickets:
                t.wait()
            torch.cuda.synchronize()
        evs = []
        for e in prof.events():
            if e.device_type is not None and "cuda" in str(e.device_type).lower():
                evs.append((e.time_range.start, e.time_range.end, e.name[:40]))
        evs.sort()
        t0 = evs[0][0] if evs else 0
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/e2e_trace.txt", "w") as f:
            for a, b, n in evs:
                f.write("%10.1f %10.1f %8.1f  %s\n" % (a - t0, b - t0, b - a, n))
        print("trace events:", len(evs))
        return
    for mode in ("full", "no-upload", "generate only"):
        tickets = []
        for k in range(4):
            tickets.append(pipe.submit(sets[k % 2], uploads[k % 2]))
        for t in tickets:
            t.wait()
        torch.cuda.synchronize()
        n = 20
        t0 = time.perf_counter()
        host = 0.0
        tickets = []
        for k in range(n):
            h0 = time.perf_counter()
            if mode == "generate only":
                ds.generate_batch(sets[k % 2])
            else:
                tickets.append(pipe.submit(sets[k % 2], uploads[k % 2] if mode == "full" else ()))
            host += time.perf_counter() - h0
            if len(tickets) > 2:
                tickets.pop(0).wait()
        for t in tickets:
            t.wait()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("%-14s %.2f ms/step wall, %.2f ms/step host inside submit -> %.0f samples/s" %
              (mode, 1e3 * dt / n, 1e3 * host / n, n * B / dt))


if __name__ == "__main__":
    main()
