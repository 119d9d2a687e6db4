"""Feasibility probe: torch symmetric memory (peer pointers over NVLink) on this box.  torchrun, 2+ ranks."""
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import torch.distributed._symmetric_memory as symm
    try:
        t = symm.empty(1 << 22, dtype=torch.float32, device=dev)
        hdl = symm.rendezvous(t, dist.group.WORLD)
        t.fill_(float(rank + 1))
        hdl.barrier()
        peer = (rank + 1) % world
        view = hdl.get_buffer(peer, (1 << 22,), torch.float32)
        got = torch.empty(1 << 22, dtype=torch.float32, device=dev)
        got.copy_(view)
        torch.cuda.synchronize()
        ok = bool((got == float(peer + 1)).all())
        # timing: barrier + peer copy of 16 MB, enqueue cost on the host and device time
        for _ in range(5):
            hdl.barrier(); got.copy_(view)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0 = time.perf_counter()
        e0.record()
        for _ in range(50):
            hdl.barrier()
            got.copy_(view)
        e1.record()
        h1 = time.perf_counter()
        torch.cuda.synchronize()
        print("rank %d: symmetric memory ok=%s, host %.1f us / (barrier + 16 MB peer copy), device %.1f us" %
              (rank, ok, 1e6 * (h1 - h0) / 50, 1e3 * e0.elapsed_time(e1) / 50), flush=True)
        # NCCL send/recv of the same size for comparison
        a = torch.empty(1 << 22, dtype=torch.float32, device=dev)
        b = torch.empty(1 << 22, dtype=torch.float32, device=dev)
        def xchg():
            ops = [dist.P2POp(dist.isend, a, (rank + 1) % world), dist.P2POp(dist.irecv, b, (rank - 1) % world)]
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        for _ in range(5):
            xchg()
        torch.cuda.synchronize()
        h0 = time.perf_counter()
        e0.record()
        for _ in range(50):
            xchg()
        e1.record()
        h1 = time.perf_counter()
        torch.cuda.synchronize()
        print("rank %d: NCCL batch_isend_irecv 16 MB: host %.1f us, device %.1f us" %
              (rank, 1e6 * (h1 - h0) / 50, 1e3 * e0.elapsed_time(e1) / 50), flush=True)
    except Exception as e:
        print("rank %d: symmetric memory FAILED: %r" % (rank, e), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
