"""Host-side multi-GPU logic on CPU: world_size-2 (and 3) gloo process groups on 127.0.0.1.

Covers what the N>1 path consists of (brainfm_b200/parallel.py): the by-sample sharding and seeding (no
collective on the data path), the max-over-ranks timing reduction of bench.py, and the plane (halo) exchange
of the 512^3 slab mode."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from brainfm_b200 import parallel as par


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    par.init(backend="gloo")
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(world, fn):
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def test_shards_are_disjoint_and_cover_everything():
    for n, world, batch in [(17, 2, 1), (64, 8, 8), (10, 4, 3), (5, 8, 1), (100, 3, 8)]:
        for epoch in range(3):
            shards = [par.shard_indices(n, r, world, epoch, batch) for r in range(world)]
            flat = sorted(i for s in shards for i in s)
            assert flat == list(range(n)), (n, world, batch, epoch)
    with pytest.raises(ValueError):
        par.shard_indices(4, 2, 2)


def test_rank_seeds_are_distinct_and_reproducible():
    seeds = {par.rank_seed(1234, r, e) for r in range(8) for e in range(16)}
    assert len(seeds) == 8 * 16
    assert par.rank_seed(7, 3, 2) == par.rank_seed(7, 3, 2)


def test_slab_bounds_partition():
    for n, world in [(512, 8), (160, 3), (7, 8), (64, 1)]:
        b = [par.slab_bounds(n, r, world) for r in range(world)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
        assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1


def _timing_fn(rank, world):
    # bench.py: every rank times its own loop, the reported time is the max over ranks
    return par.all_reduce_max(10.0 + rank)


def test_max_over_ranks_gloo():
    assert _run(2, _timing_fn) == [11.0, 11.0]


def _halo_fn(rank, world):
    n, halo_lo, halo_hi = 23, 3, 2
    full = torch.arange(n * 4 * 5, dtype=torch.float32).reshape(n, 4, 5)
    owned = [list(par.slab_bounds(n, r, world)) for r in range(world)]
    needed = [[max(0, b - halo_lo), min(n, e + halo_hi)] for b, e in owned]
    local = full[owned[rank][0]:owned[rank][1]].clone()
    got = par.exchange_planes(local, owned, needed)
    ok = torch.equal(got, full[needed[rank][0]:needed[rank][1]])
    # a second, asymmetric pattern: everybody needs the first two planes (owned by rank 0) plus its own
    needed2 = [[0, e] if r > 0 else [0, owned[0][1]] for r, (b, e) in enumerate(owned)]
    needed2 = [[0, 2]] * world
    got2 = par.exchange_planes(local, owned, needed2)
    ok2 = torch.equal(got2, full[0:2])
    return bool(ok and ok2)


@pytest.mark.parametrize("world", [2, 3])
def test_plane_exchange_gloo(world):
    assert all(_run(world, _halo_fn))


def _by_sample_fn(rank, world):
    # the data path of the sample-parallel mode: disjoint indices, disjoint seeds, NO collective;
    # only the final timing reduction touches the process group
    idx = par.shard_indices(32, rank, world, epoch=1, batch=8)
    np.random.seed(par.rank_seed(1000, rank))
    draw = float(np.random.rand())
    t = par.all_reduce_max(float(len(idx)))
    return (idx, draw, t)


def test_by_sample_mode_gloo():
    out = _run(2, _by_sample_fn)
    assert sorted(out[0][0] + out[1][0]) == list(range(32))
    assert out[0][1] != out[1][1]
    assert out[0][2] == out[1][2] == 16.0


def test_slab_host_ranges_match_the_band_and_zoom_tables():
    """Generator/slab.py plans its plane exchanges on the host from cached per-axis ranges: the first tap of the banded
    x pass must be band_host's `start`, the zoom-back ranges the zoom tables' lo / hi -- for every (n_in, n_out) a 64^3
    or 160^3 volume can draw."""
    import numpy as np
    from brainfm_b200.Generator import slab
    from brainfm_b200.plan import band_host, zoom_tables_host
    for n_in, n_out, sigma in [(64, 64, 0.0), (64, 23, 1.7), (160, 160, 0.9), (160, 53, 2.4), (160, 18, 4.6), (512, 171, 2.1)]:
        lo, centre = slab._band_ranges_host(n_in, n_out)
        start, w, T = band_host(n_in, n_out, sigma)
        assert np.array_equal(lo - (T - 2) // 2, start)
        assert centre.min() >= 0 and centre.max() <= n_in - 1 and np.all(np.diff(centre) >= 0)
        zlo, zhi = slab._zoom_ranges_host(n_out, n_in)
        t = zoom_tables_host(n_out, 1 / (n_out / n_in), n_in)
        assert np.array_equal(zlo, t[0]) and np.array_equal(zhi, t[1])
        assert np.all(np.diff(zlo) >= 0) and zhi.max() <= n_out - 1
    # every plane a rank's low-res outputs read lies inside the range the exchange asks for
    n_in, n_out, sigma, world = 160, 41, 3.3, 4
    lo, centre = slab._band_ranges_host(n_in, n_out)
    start, w, T = band_host(n_in, n_out, sigma)
    from brainfm_b200 import parallel as par
    for r in range(world):
        b, e = par.slab_bounds(n_in, r, world)
        o = np.nonzero((centre >= b) & (centre < e))[0]
        if o.size:
            need = (max(0, start[o[0]:o[-1] + 1].min()), min(n_in, start[o[0]:o[-1] + 1].max() + T))
            taps = np.concatenate([np.arange(max(0, s), min(n_in, s + T)) for s in start[o[0]:o[-1] + 1]])
            assert taps.min() >= need[0] and taps.max() < need[1]
