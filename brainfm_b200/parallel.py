"""Multi-GPU plumbing of the generator: one process per GPU, `torch.distributed` (NCCL over NVLink on GPUs,
gloo in the CPU tests).

The reference generator issues no collective: every `__getitem__` is independent, so the path
shards BY SAMPLE -- disjoint index ranges and disjoint random streams per rank, nothing exchanged
(`shard_indices`, `rank_seed`).  The one case with a real exchange step is a single volume too large for the
per-sample working set to be worth replicating (512^3): it is cut into x-slabs of the OUTPUT grid and the two
stencil stages (slice-profile blur along x, low-res -> training-grid zoom along x) read a few planes owned by
the neighbouring ranks (`slab_bounds`, `exchange_planes`); two scalars are all-reduced (`all_reduce_max`).
"""
import os

import torch
import torch.distributed as dist


def env_rank():
    """(rank, world, local_rank) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend=None):
    """Initialise the default process group when WORLD_SIZE > 1 (idempotent).  NCCL when CUDA is available,
    else gloo; the rendezvous defaults to 127.0.0.1 (single node)."""
    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local


# ------------------------------------------------------------------------------------------------ by sample
def shard_indices(n_items, rank, world, epoch=0, batch=1):
    """Indices of the items rank `rank` generates in one epoch: whole batches dealt round-robin, rotated by the
    epoch so that every rank sees every subject over `world` epochs.  The shards of all ranks are disjoint and
    together cover range(n_items) exactly once."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    n_batches = (n_items + batch - 1) // batch
    mine = [b for b in range(n_batches) if (b + epoch) % world == rank]
    return [i for b in mine for i in range(b * batch, min((b + 1) * batch, n_items))]


def rank_seed(base, rank, epoch=0):
    """Distinct, reproducible generator seed per (base, rank, epoch): splitmix64 of the triple."""
    z = (int(base) * 0x9E3779B97F4A7C15 + (rank + 1) * 0xBF58476D1CE4E5B9 + (epoch + 1) * 0x94D049BB133111EB)
    z &= 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return int((z ^ (z >> 31)) & 0x7FFFFFFF)


def all_reduce_max(value, device=None):
    """Maximum over ranks of a python float (timings are reported as the max over ranks) or of a tensor."""
    if isinstance(value, torch.Tensor):
        t = value.clone()
    else:
        t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t if isinstance(value, torch.Tensor) else float(t.item())


# ------------------------------------------------------------------------------------------------ by slab
def slab_bounds(n, rank, world):
    """[begin, end) of the planes rank `rank` owns when n planes are cut into `world` near-equal slabs."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def exchange_planes(local, owned, needed, rank=None, world=None, group=None):
    """Halo exchange along axis 0.

    local  : tensor holding this rank's planes [owned[rank][0], owned[rank][1]) (axis 0)
    owned  : list of [begin, end) per rank -- disjoint, known to every rank (no negotiation round)
    needed : list of [begin, end) per rank -- the planes each rank must see (a superset of what it owns is
             not required; planes nobody owns must not be asked for)
    Returns a tensor with planes [needed[rank][0], needed[rank][1]).  Each rank sends a peer exactly the planes
    it owns and the peer needs; transfers are point-to-point (`batch_isend_irecv`: NCCL send/recv over NVLink
    on GPUs), there is no collective and no staging through the host."""
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    ob, oe = owned[rank]
    nb, ne = needed[rank]
    if local.shape[0] != oe - ob:
        raise ValueError("local tensor has %d planes, rank owns %d" % (local.shape[0], oe - ob))
    if world > 1 and local.is_cuda and dist.get_backend(group) == "gloo":
        # gloo moves host memory only (several ranks sharing one GPU in the tests): stage the planes through the host
        host = exchange_planes(local.cpu(), owned, needed, rank, world, group)
        return host.to(local.device)
    out = torch.empty((max(ne - nb, 0), *local.shape[1:]), dtype=local.dtype, device=local.device)
    # own planes
    a, b = max(ob, nb), min(oe, ne)
    if b > a:
        out[a - nb:b - nb].copy_(local[a - ob:b - ob])
    ops, keep = [], []
    for peer in range(world):
        if peer == rank:
            continue
        pb, pe = owned[peer]
        qb, qe = needed[peer]
        # what I send: my planes the peer needs
        a, b = max(ob, qb), min(oe, qe)
        if b > a:
            t = local[a - ob:b - ob].contiguous()
            keep.append(t)
            ops.append(dist.P2POp(dist.isend, t, peer, group))
        # what I receive: the peer's planes I need
        a, b = max(pb, nb), min(pe, ne)
        if b > a:
            ops.append(dist.P2POp(dist.irecv, out[a - nb:b - nb], peer, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out
