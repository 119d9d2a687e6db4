"""torchrun worker: slab mode WITHOUT injected draws.  Every rank generates its x-slab with the library's own
counter-based noise; rank 0 checks that the assembled volume equals the single-rank slab run bit for bit (the noise
is keyed on absolute voxels, not on the decomposition) and agrees with the fused chain (generate_batch) for the same
seeds within the float tolerance.  Ranks may share one GPU (BFM_SLAB_ONE_GPU=1, gloo backend: CUDA tensors are staged
through the host), which is how the 1-GPU test box runs it.  Exit code 0 = pass.

    python -m torch.distributed.run --nproc-per-node 3 --master-addr 127.0.0.1 tests/_slab_noise_worker.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import bench
from brainfm_b200 import parallel as par
from brainfm_b200.Generator.slab import generate_slab

SIZE = 64


def dataset(dev):
    from brainfm_b200 import io as bio
    bio.clear_registry()
    old = bench.SIZE
    bench.SIZE = SIZE
    try:
        ds = bench.build_dataset(bench.make_inputs(1), dev, planner=os.environ.get("BFM_SLAB_PLANNER", "python"))
    finally:
        bench.SIZE = old
    return ds


def seeded(fn, seed, ds=None):
    if ds is not None:                       # the Philox key of a sample = splitmix64(one numpy draw, call counter):
        ds.rng._seed_base, ds.rng._seed_count = None, 0        # restart it so that equal seeds give equal keys
        if getattr(ds, "_native", None) is not None:             # library planner: (seed, item counter) likewise
            ds._native.seed, ds._native.counter = None, 0
    np.random.seed(seed)
    torch.manual_seed(seed)
    import random
    random.seed(seed)
    return fn()


def main():
    one_gpu = os.environ.get("BFM_SLAB_ONE_GPU") == "1"
    rank, world, local = par.init("gloo" if one_gpu else None)
    dev = torch.device("cuda", 0 if one_gpu else local)
    torch.cuda.set_device(dev)
    ok = True
    for seed in (3, 11, 12):                          # different resolution classes / flips
        ds = dataset(dev)
        mine = seeded(lambda: generate_slab(ds, 0, rank, world), seed, ds)
        assert ds._last_descs[0][0].eps_noise is None and ds._last_descs[0][0].eps_gmm is None
        if os.environ.get("BFM_SLAB_PLANNER") == "native":
            assert ds._native is not None and ds._last_descs[0][0].gen_small != 0      # planned by the library
        solo = seeded(lambda: generate_slab(ds, 0, 0, 1), seed, ds)
        item = seeded(lambda: ds.generate_batch([0])[0], seed, ds)
        fused = item[4]['input']
        torch.cuda.synchronize()
        x0, x1 = mine["x_range"]
        same = torch.equal(mine["input"], solo["input"][:, x0:x1])
        close = np.allclose(solo["input"].cpu().numpy(), fused.cpu().numpy(), rtol=1e-5, atol=1e-4)
        if os.environ.get("BFM_SLAB_PLANNER") == "native":
            # the real-image target that rides on the gather is sharded too: every rank's planes equal the single-rank
            # run's and the fused chain's target bit for bit (same gathers, the volume's min / max all-reduced)
            assert "T1" in mine and torch.is_tensor(item[3]["T1"])
            same = same and torch.equal(mine["T1"], solo["T1"][:, x0:x1]) and \
                torch.equal(solo["T1"], item[3]["T1"])
        if "bias_field_log" in mine:
            same = same and torch.equal(mine["bias_field_log"], solo["bias_field_log"][:, x0:x1])
        if rank == 0:
            print("seed %d: rank slab == single-rank slab: %s; slab mode vs fused chain within tolerance: %s "
                  "(max |diff| %.3g)" % (seed, same, close, float((solo["input"] - fused).abs().max())))
        ok = ok and same and close
    flag = torch.tensor([1 if ok else 0])
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
