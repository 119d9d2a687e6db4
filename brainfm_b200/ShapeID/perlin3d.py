"""Perlin / fractal noise, random shapes and random divergence-free velocities
(mirror of ShapeID/perlin3d.py).  The noise kernel reproduces the numpy float64 reference bit for bit; the random
lattice gradients are drawn on the host from numpy's global generator exactly like the reference."""
import numpy as np
import torch

from .. import _lib
from ._common import ivec, stream
from .misc import stream_3D


def interpolant(t):
    return t * t * t * (t * (t * 6 - 15) + 10)


def _lattice(res, tileable, draws=None):
    shp = (res[0] + 1, res[1] + 1, res[2] + 1)
    if draws is None:
        theta = 2 * np.pi * np.random.rand(*shp)
        phi = 2 * np.pi * np.random.rand(*shp)
    else:                               # injectable draws (brainfm_b200.draws), same order
        theta = 2 * np.pi * draws.rand_array("perlin.theta", shp)
        phi = 2 * np.pi * draws.rand_array("perlin.phi", shp)
    g = np.stack((np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta), np.cos(phi)), axis=3)
    if tileable[0]:
        g[-1, :, :] = g[0, :, :]
    if tileable[1]:
        g[:, -1, :] = g[:, 0, :]
    if tileable[2]:
        g[:, :, -1] = g[:, :, 0]
    return g


def _percentile_device(noise, percentile):
    """np.percentile(noise, q) (linear interpolation between order statistics) from device-side order
    statistics; the two neighbouring order statistics are exact, the lerp follows numpy's _lerp."""
    flat = noise.reshape(-1)
    n = flat.numel()
    virt = percentile / 100.0 * (n - 1)
    lo = int(np.floor(virt))
    hi = min(lo + 1, n - 1)
    g = virt - lo
    vals = torch.sort(flat).values[[lo, hi]].cpu().numpy()
    a, b = float(vals[0]), float(vals[1])
    d = b - a
    out = a + d * g
    if g >= 0.5:
        out = b - d * (1 - g)
    if d == 0:
        out = a
    return out


def _noise_device(shape, res, tileable, device, draws=None):
    shape, res = [int(s) for s in shape], [int(r) for r in res]
    for s, r in zip(shape, res):
        if s % r:
            raise ValueError("shape must be a multiple of res")
    g = torch.from_numpy(np.ascontiguousarray(_lattice(res, tileable, draws))).to(device)
    out = torch.empty(shape, dtype=torch.float64, device=device)
    _lib.check(_lib.lib().bfm_perlin3d(g.data_ptr(), ivec(shape), ivec(res), out.data_ptr(), stream()))
    return out


def _finish(noise, percentile, as_numpy):
    if percentile is None:
        return noise.cpu().numpy() if as_numpy else noise
    thr = _percentile_device(noise, percentile)
    mask = torch.empty_like(noise)
    _lib.check(_lib.lib().bfm_threshold_mask(noise.data_ptr(), mask.data_ptr(), noise.numel(), float(thr), stream()))
    if as_numpy:
        return noise.cpu().numpy(), mask.cpu().numpy()
    return noise, mask


def generate_perlin_noise_3d(shape, res, tileable=(False, False, False), interpolant=interpolant, percentile=None,
                             device='cuda', as_numpy=True, draws=None):
    """3-D Perlin noise (ShapeID/perlin3d.py:15-90).  Returns numpy arrays like the reference unless
    as_numpy=False (device tensors, no D2H)."""
    noise = _noise_device(shape, res, tileable, torch.device(device), draws)
    return _finish(noise, percentile, as_numpy)


def generate_fractal_noise_3d(shape, res, octaves=1, persistence=0.5, lacunarity=2, tileable=(False, False, False),
                              interpolant=interpolant, percentile=None, device='cuda', as_numpy=True):
    """Sum of Perlin octaves (ShapeID/perlin3d.py:94-141)."""
    noise = torch.zeros([int(s) for s in shape], dtype=torch.float64, device=device)
    frequency, amplitude = 1, 1
    for _ in range(octaves):
        noise += amplitude * _noise_device(shape, (frequency * res[0], frequency * res[1], frequency * res[2]),
                                           tileable, torch.device(device))
        frequency *= lacunarity
        amplitude *= persistence
    return _finish(noise, percentile, as_numpy)


def generate_shape_3d(shape, perlin_res, percentile, device, draws=None):
    """Random blob: (mask, noise*mask) as float64 device tensors (ShapeID/perlin3d.py:144-146)."""
    pprob, p = generate_perlin_noise_3d(shape, perlin_res, tileable=(True, False, False), percentile=percentile,
                                        device=device, as_numpy=False, draws=draws)
    return p, pprob


def generate_velocity_3d(shape, perlin_res, V_multiplier, device, draws=None):
    """Divergence-free velocity: curl of three Perlin potentials, float32 (ShapeID/perlin3d.py:149-156)."""
    dev = torch.device(device)
    a = _noise_device(shape, perlin_res, (True, False, False), dev, draws)
    b = _noise_device(shape, perlin_res, (True, False, False), dev, draws)
    c = _noise_device(shape, perlin_res, (True, False, False), dev, draws)
    Vx, Vy, Vz = stream_3D(a, b, c, multiplier=V_multiplier)
    return {'Vx': Vx, 'Vy': Vy, 'Vz': Vz}
