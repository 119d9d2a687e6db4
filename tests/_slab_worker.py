"""torchrun worker of the multi-GPU slab test: every rank generates its x-slab of one oracle case with the
oracle's draws injected; rank 0 assembles the volume and compares it with the oracle and with the slab-mode
result of a single rank.  Exit code 0 = pass.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/_slab_worker.py g64_s0
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

from brainfm_b200 import parallel as par
from brainfm_b200.Generator.slab import generate_slab
from oracle import make_golden as mg
from tests._harness import cuda_case, oracle_case, to_np


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "g64_s0"
    rank, world, local = par.init()
    torch.cuda.set_device(local)
    item, orc = oracle_case(name)
    ref = mg.flatten(item)
    mine, ds, draws = cuda_case(name, orc.log, run=lambda ds: generate_slab(ds, 0, rank, world))
    solo, _, _ = cuda_case(name, orc.log, run=lambda ds: generate_slab(ds, 0, 0, 1))
    torch.cuda.synchronize()
    ok = True
    for key, rkey in (("input", "sample0.input"), ("bias_field_log", "sample0.bias_field_log")):
        if key not in mine:
            continue
        x0, x1 = mine["x_range"]
        # bit-identical to the single-rank slab run on the owned planes
        same = torch.equal(mine[key], solo[key][:, x0:x1])
        ranges = _ranges(ds, world)
        parts = [torch.empty_like(solo[key][:, b:e]).contiguous() for b, e in ranges]
        if world > 1:
            # planes per rank may differ by one: pad to the largest slab for the collective
            most = max(e - b for b, e in ranges)
            pad = torch.zeros((1, most, *mine[key].shape[2:]), device="cuda")
            pad[:, :x1 - x0] = mine[key]
            bufs = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(bufs, pad)
            parts = [bufs[r][:, :e - b] for r, (b, e) in enumerate(ranges)]
        else:
            parts = [mine[key]]
        order = sorted(range(world), key=lambda r: ranges[r][0])
        full = torch.cat([parts[r] for r in order], dim=1)
        if rank == 0:
            a, b = to_np(ref[rkey]), to_np(full)
            err = float(np.abs(a - b).max())
            close = np.allclose(b, a, rtol=1e-5, atol=1e-4)
            print("%s: slab==solo %s, max |slab - oracle| = %.3g, within tolerance %s" % (key, same, err, close))
            ok = ok and close
        ok = ok and same
    flag = torch.tensor([1 if ok else 0], device="cuda")
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


def _ranges(ds, world):
    """Output plane ranges of all ranks, in rank order (all_gather needs the shapes up front)."""
    s0 = int(ds.size[0])
    flip = bool(ds.last_setups['flip'])
    out = []
    for r in range(world):
        b, e = par.slab_bounds(s0, (world - 1 - r) if flip else r, world)
        out.append((s0 - e, s0 - b) if flip else (b, e))
    return out


if __name__ == "__main__":
    main()
