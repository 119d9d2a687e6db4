"""One large volume across several GPUs: x-slabs of the output grid with plane (halo) exchange.

The reference has no spatial decomposition; its generator simply cannot hold the
intermediates of a 512^3 sample next to a training job.  Here the OUTPUT grid is cut into x-slabs, one per
rank; every rank holds the (integer) source label map and draws the same random numbers (same seeds), so the
deformation, the contrast tables and the GMM image are identical everywhere without any traffic (the GMM noise
is counter-based, keyed on the absolute source voxel).  Only two stages have a stencil along x and therefore
an exchange step, both point-to-point between neighbouring ranks (parallel.exchange_planes -> NCCL send/recv):

  1. the slice-profile blur + downsample along x needs ceil(3*sigma)+1 planes of the warped image from each
     neighbour (<= 17 planes of 512^2 floats = 17.8 MB per side);
  2. the zoom back to the training grid needs the 1-2 low-res planes next to the slab.

The global maximum of I / max(I) is one scalar all-reduce.  Per voxel the arithmetic is that of the op-level
path (resample_resolution -> add_noise -> myzoom_torch) and both volume-sized noise fields are counter-based (GMM
noise keyed on the absolute source voxel, acquisition noise on the absolute low-res voxel, bfm_add_noise_at), so a
volume generated on W ranks is bit-identical to the same volume generated in slab mode on one rank -- with injected
draws or without (tests/test_gen_parity_gpu.py, tests/_slab_worker.py).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from .. import _lib, parallel as par
from ..plan import band_host, resample_coords_host, zoom_tables_host
from .utils import _stream


def _band_axis(x, axis, n_out, start, w, T, arena):
    # tables through the pinned plan arena (asynchronous upload): a pageable host->device copy would wait for all
    # queued GPU work every time and serialise host and device
    p_start = arena.put(np.ascontiguousarray(start, dtype=np.int32))
    p_w = arena.put(np.ascontiguousarray(w, dtype=np.float32))
    arena.commit()
    shape = list(x.shape)
    out_shape = list(shape)
    out_shape[axis] = int(n_out)
    y = torch.empty(out_shape, dtype=torch.float32, device=x.device)
    if y.numel() and x.numel():
        _lib.check(_lib.lib().bfm_band_axis(x.data_ptr(), y.data_ptr(), (C.c_int * 3)(*shape), axis, int(n_out),
                                            p_start, p_w, int(T), -1.0, None, 0, _stream()))
    return y


def generate_slab(ds, idx, rank=None, world=None, group=None):
    """This rank's x-slab of sample `idx` of dataset `ds` (BaseGen): {'input': (1, nx, s1, s2),
    'bias_field_log': (1, nx, s1, s2) or absent, 'x_range': (x0, x1)} with x0:x1 the owned planes of the final
    (flipped) output volume.  Every rank must call this with the same generator seeds."""
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    L = _lib.lib()
    size = [int(v) for v in ds.size]
    s0, s1, s2 = size
    arena = ds.arena.begin()
    ctx = ds._prologue_host(idx, arena)
    if not ds._fast_ok(ctx['input_mode']):
        raise NotImplementedError("slab mode covers synthetic inputs with the stock augmentation chain")
    setups = ctx['setups']
    target = {}
    for a in ds._gen_arg_sets()[0]:
        ds.update_gen_args(a)
    p = ds._plan_synth(setups, target)
    if p['mix'] is not None:
        raise NotImplementedError("slab mode does not mix real modalities into the synthetic image")
    job = ds._job(setups, ctx['deform'], target, p)
    descs, results = ds._build_descs([job], arena)
    flip = bool(setups['flip'])
    # planes of the (unflipped) grid this rank computes; with a flip the slabs are mirrored so that the planes
    # a rank computes are the ones it owns in the flipped output
    owned = [list(par.slab_bounds(s0, (world - 1 - r) if flip else r, world)) for r in range(world)]
    c0, c1 = owned[rank]
    descs[0].x_begin, descs[0].x_count = c0, c1 - c0
    d_dev = arena.put_struct_array(descs)
    arena.commit()
    ds._last_descs = (descs, d_dev, 1)
    h, st = C.addressof(descs), _stream()
    _lib.check(L.bfm_gen_bbox(h, d_dev, 1, st))
    job['plan'].have_bbox = True
    _lib.check(L.bfm_gen_gmm(h, d_dev, 1, st))
    _lib.check(L.bfm_gen_warp(h, d_dev, 1, st))
    N = s0 * s1 * s2
    i_bf = ds._ws['i_bf'][:N].view(s0, s1, s2)[c0:c1]

    # ---- resolution degradation: x first (needs neighbours' planes), then y and z on the slab
    new = [int(v) for v in p['new_size']]
    stds = [float(v) for v in p['stds']]
    start, w, T = band_host(s0, new[0], stds[0])
    centre = np.clip(np.floor(resample_coords_host(s0, new[0])[0]).astype(np.int64), 0, s0 - 1)
    low_owned = []
    for r in range(world):
        b, e = owned[r]
        o = np.nonzero((centre >= b) & (centre < e))[0]
        low_owned.append([int(o[0]), int(o[-1]) + 1] if o.size else [0, 0])
    needed = []
    for r in range(world):
        ob, oe = low_owned[r]
        if oe > ob:
            needed.append([int(max(0, start[ob:oe].min())), int(min(s0, start[ob:oe].max() + T))])
        else:
            needed.append([owned[r][0], owned[r][0]])
    ext = par.exchange_planes(i_bf, owned, needed, rank, world, group)
    ob, oe = low_owned[rank]
    x = _band_axis(ext, 0, oe - ob, start[ob:oe] - needed[rank][0], w[ob:oe], T, arena)
    for ax in (1, 2):
        st_a, w_a, T_a = band_host(size[ax], new[ax], stds[ax])
        x = _band_axis(x, ax, new[ax], st_a, w_a, T_a, arena)
    # ---- noise (add_noise, utils.py:633-638): injected draws are sliced; otherwise the chain's counter-based stream 1,
    # keyed on the ABSOLUTE low-res voxel -- the assembled volume does not depend on the number of ranks, and equals
    # what the fused chain (generate_batch) draws for the same seed
    if p['eps_noise'] is not None:
        eps = p['eps_noise'][ob:oe].to(x.device)
        low = torch.clamp(x + float(p['noise_std']) * eps, min=0)
    else:
        low = x.contiguous()
        _lib.check(L.bfm_add_noise_at(low.data_ptr(), low.numel(), float(p['noise_std']), int(p['seed']), 1,
                                      ob * new[1] * new[2], _stream()))

    # ---- back to the training grid: low-res planes lo[c0] .. hi[c1-1] (1-2 from the neighbours)
    up = 1 / (np.array(new) / np.array(size))
    tabs = [zoom_tables_host(new[a], up[a], size[a]) for a in range(3)]
    need_low = []
    for r in range(world):
        b, e = owned[r]
        need_low.append([int(tabs[0][0][b:e].min()), int(tabs[0][1][b:e].max()) + 1] if e > b else [0, 0])
    lext = par.exchange_planes(low, low_owned, need_low, rank, world, group)
    nb = need_low[rank][0]
    t0 = (tabs[0][0][c0:c1] - nb, tabs[0][1][c0:c1] - nb, tabs[0][2][c0:c1], tabs[0][3][c0:c1])
    flat = [np.ascontiguousarray(t) for t in t0] + [t for tab in tabs[1:] for t in tab]
    addr = [arena.put(np.ascontiguousarray(t)) for t in flat]
    arena.commit()
    out = torch.empty((c1 - c0, s1, s2), dtype=torch.float32, device=low.device)
    if out.numel():
        args = [lext.data_ptr(), lext.shape[0], new[1], new[2], 1]
        for d, n_out in enumerate((c1 - c0, s1, s2)):
            args += [addr[4 * d], addr[4 * d + 1], addr[4 * d + 2], addr[4 * d + 3], int(n_out)]
        _lib.check(L.bfm_zoom_linear(*args, out.data_ptr(), _stream()))
    # ---- I / max(I) with the global maximum (datasets.py:342-343), then the flip
    mx = out.max().reshape(1) if out.numel() else torch.zeros(1, device=low.device)
    if world > 1:
        if dist.get_backend(group) == "gloo":              # ranks sharing one GPU (tests): reduce on the host
            h_mx = mx.cpu()
            dist.all_reduce(h_mx, op=dist.ReduceOp.MAX, group=group)
            mx = h_mx.to(mx.device)
        else:
            dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    # one pass: I / max and the flip of the slab's planes (bfm_shift_scale_flip: true division like the reference)
    if out.numel():
        fin = torch.empty_like(out)
        _lib.check(L.bfm_shift_scale_flip(out.data_ptr(), fin.data_ptr(), c1 - c0, s1 * s2, None, mx.data_ptr(), 1.0,
                                          1 if flip else 0, _stream()))
        out = fin
    sample = {}
    if flip:
        x0, x1 = s0 - c1, s0 - c0
    else:
        x0, x1 = c0, c1
    sample['input'] = out[None]
    if 'bias_field_log' in results[0]:
        sample['bias_field_log'] = results[0]['bias_field_log'][:, x0:x1]
    sample['x_range'] = (x0, x1)
    arena.mark_done()
    ds.last_setups, ds.last_deform = setups, ctx['deform']
    return sample
