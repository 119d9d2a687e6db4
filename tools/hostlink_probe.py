"""Host <-> device copy ceiling of the box, all GPUs at once.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/hostlink_probe.py

Every rank pins 2 x 256 MB of host memory (first touch after the optional NUMA binding), then times plain
cudaMemcpyAsync traffic on its own GPU -- H2D alone, D2H alone, both directions at once, and the e2e arm's own mix
(98 MB in + 131 MB out per step) -- with all ranks running concurrently between barriers.  Rank 0 prints one JSON
line with the per-GPU and aggregate GB/s; run once with BFM_NUMA_BIND=0 and once with =1 to see what binding the
process (and therefore its pinned buffers) to the GPU's NUMA node buys.  No kernels of the library are involved:
this is the denominator of the end-to-end arm, not a bench value."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    bind = os.environ.get("BFM_NUMA_BIND", "1") != "0"
    node = bench.bind_to_gpu_numa_node(local) if bind else None
    if world > 1:
        dist.init_process_group("gloo")
    dev = torch.device("cuda", local)
    MB = 1 << 20
    n_in, n_out = 256 * MB, 256 * MB
    h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    h_out.fill_(2)
    d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n_out, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def run(up_bytes, down_bytes, reps=8):
        for timed in (False, True):
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps if timed else 2):
                if up_bytes:
                    with torch.cuda.stream(s1):
                        d_in[:up_bytes].copy_(h_in[:up_bytes], non_blocking=True)
                if down_bytes:
                    with torch.cuda.stream(s2):
                        h_out[:down_bytes].copy_(d_out[:down_bytes], non_blocking=True)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / reps
        t = torch.tensor([dt], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    out = {"n_gpus": world, "numa_bind": bind, "numa_node_rank0": node, "cpus": len(os.sched_getaffinity(0))}
    t = run(n_in, 0)
    out["h2d_gbs_per_gpu"] = n_in / t / 1e9
    t = run(0, n_out)
    out["d2h_gbs_per_gpu"] = n_out / t / 1e9
    t = run(n_in, n_out)
    out["bidir_gbs_per_gpu_each_way"] = n_in / t / 1e9
    up, down = 8 * 160 ** 3 * 3, 8 * 160 ** 3 * 4          # one e2e step: u8 labels + int16 T1 in, f32 input out
    t = run(up, down, reps=16)
    out["e2e_mix_ms_per_step"] = t * 1e3
    out["e2e_mix_ceiling_samples_per_s"] = world * 8 / t
    for k in ("h2d_gbs_per_gpu", "d2h_gbs_per_gpu", "bidir_gbs_per_gpu_each_way"):
        out[k.replace("per_gpu", "aggregate")] = out[k] * world
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
