"""512^3 volume generated in slab mode on N GPUs (BASELINE.json configs[4], second half): x-slabs of the output
grid, plane exchange over NCCL for the x stencils.  Launch with torchrun for N > 1:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/slab_bench.py

One JSON line on rank 0: ms per volume (max over ranks, CUDA events), volumes/s, voxels/s."""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from brainfm_b200 import io as bio, parallel as par
from brainfm_b200.Generator import BaseGen
from brainfm_b200.Generator.slab import generate_slab
from tests import _inputs as ti


def run(size=512, steps=5, rank=0, world=1, local=0):
    """ms per volume (max over ranks) of one size^3 sample in slab mode on `world` ranks + a checksum of the assembled
    volume (sum over ranks of the owned planes' sums: equal for every decomposition)."""
    dev = torch.device("cuda", local)
    half = ti.brain_like_labels((size // 2,) * 3, seed=7)
    lab = np.repeat(np.repeat(np.repeat(half, 2, 0), 2, 1), 2, 2)       # nearest x2
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
    ds = BaseGen(cfg, dev)
    ds.write_bflog = True

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    np.random.seed(4321)
    torch.manual_seed(4321)                  # the same draws on every rank
    # volumes in flight: each on its own stream with its own scratch (DevicePipeline), so that one volume's rendezvous
    # waits (two plane exchanges, one all-reduce) are filled with the other's kernels.  Every rank issues the same
    # sequence of NCCL operations, whatever the lane.
    # (measured, same 20 volumes: 2.43 -> 2.28 ms on 1 GPU, 1.47 -> 1.31 on 2, 0.89 on 4; on 8 GPUs the host side -- 8
    # Python processes and their NCCL proxy threads on the box's 32 vCPUs -- is the limit and a second volume in flight
    # costs more host time than it hides: 1.32 ms against 0.96 with one)
    lanes = int(os.environ.get("SLAB_LANES", "2" if world <= 4 else "1"))
    from brainfm_b200.pipeline import DevicePipeline
    pipe = DevicePipeline(ds, depth=lanes)
    job = lambda: generate_slab(ds, 0, rank, world)

    def loop(n):
        tickets = []
        for _ in range(n):
            tickets.append(pipe.submit(call=job))
            if len(tickets) > lanes:
                tickets.pop(0).wait()
        res = None
        for t in tickets:
            res = t.wait()
        return res

    loop(2 * lanes)
    sync()
    # the timed volumes are the same whatever the warm-up drew (their cost varies with the resolution class)
    np.random.seed(777)
    torch.manual_seed(777)
    if getattr(ds, "_native", None) is not None:
        ds._native.seed, ds._native.counter = None, 0
    ds.rng._seed_base, ds.rng._seed_count = None, 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    h0 = time.perf_counter()
    out = loop(steps)
    host_ms = 1e3 * (time.perf_counter() - h0) / steps          # time to ENQUEUE a volume (incl. waits for arena slots)
    e1.record()
    sync()
    ms = par.all_reduce_max(e0.elapsed_time(e1) / steps, device=dev)
    prof_txt = None
    if os.environ.get("SLAB_PROFILE") == "1":
        # rank 0: how busy is the GPU?  (sum of kernel / memcpy durations over the wall time of `steps` volumes)
        from torch.profiler import profile, ProfilerActivity
        sync()
        w0 = time.perf_counter()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(steps):
                generate_slab(ds, 0, rank, world)
            torch.cuda.synchronize()
        wall = time.perf_counter() - w0
        if rank == 0:
            ev = [e for e in prof.key_averages()]
            tot = sum(e.device_time_total for e in ev) * 1e-3
            rows = sorted(ev, key=lambda e: -e.device_time_total)[:14]
            prof_txt = "GPU busy %.3f ms of %.3f ms wall per volume (rank 0, under the profiler)\n" % (tot / steps, 1e3 * wall / steps)
            prof_txt += "\n".join("%8.1f us x%-3d %s" % (e.device_time_total / steps, e.count // steps, e.key[:90]) for e in rows)
        sync()
    chk = out["input"].double().sum().reshape(1)
    if world > 1:
        dist.all_reduce(chk)
    bio.clear_registry()
    return {"workload": "one %d^3 BaseGen sample (input + bias_field_log), slab mode" % size,
            "n_gpus": world, "ms_per_volume": ms, "volumes_per_s": 1e3 / ms,
            "Mvoxels_per_s": size ** 3 / ms / 1e3, "slab_planes_rank0": list(out["x_range"]),
            "checksum": float(chk.item()), "steps": steps,
            "host_enqueue_ms_per_volume": par.all_reduce_max(host_ms, device=dev), "volumes_in_flight": lanes,
            "profile": prof_txt}


def main():
    size = int(os.environ.get("SLAB_SIZE", "512"))
    steps = int(os.environ.get("SLAB_STEPS", "5"))
    rank, world, local = par.init()
    torch.cuda.set_device(local)
    res = run(size, steps, rank, world, local)
    if rank == 0:
        prof_txt = res.pop("profile", None)
        print(json.dumps(res), flush=True)
        if prof_txt:
            print(prof_txt, file=sys.stderr, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
