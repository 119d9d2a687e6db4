"""Resamplers next to the generator path that the reference keeps in utils/misc.py (inference pre-processing):
`myzoom_torch_anisotropic` (utils/misc.py:1051-1115) and `torch_resize` (utils/misc.py:1117-1187), on libbfm's
separable zoom (bfm_zoom_linear) and zero-padded Gaussian blur (bfm_blur_axis) kernels.  CUDA tensors only."""
import numpy as np
import torch

from . import _lib
from .Generator.utils import _need_cuda, _stream, pack_to_device
from .plan import zoom_tables_host


def _zoom_to(X4, newsize):
    a, b, c, Cn = X4.shape
    shape = (a, b, c)
    factors = np.array(newsize, dtype=np.float64) / np.array(shape, dtype=np.float64)
    tabs = [zoom_tables_host(shape[d], factors[d], int(newsize[d])) for d in range(3)]
    keep, addr = pack_to_device([t for tab in tabs for t in tab], X4.device)
    out = torch.empty((int(newsize[0]), int(newsize[1]), int(newsize[2]), Cn), dtype=torch.float32, device=X4.device)
    args = [X4.data_ptr(), a, b, c, Cn]
    for d in range(3):
        args += [addr[4 * d], addr[4 * d + 1], addr[4 * d + 2], addr[4 * d + 3], int(newsize[d])]
    _lib.check(_lib.lib().bfm_zoom_linear(*args, out.data_ptr(), _stream()))
    return out, factors


def _new_affine(aff, factors):
    aff_new = aff.copy()
    for c in range(3):
        aff_new[:-1, c] = aff_new[:-1, c] / factors[c]
    aff_new[:-1, -1] = aff_new[:-1, -1] - aff[:-1, :-1] @ (0.5 - 0.5 / factors)
    return aff_new


def myzoom_torch_anisotropic(X, aff, newsize):
    """Separable linear zoom of a (X, Y, Z[, C]) volume to `newsize` with edge clamp; returns (Y, new affine) or Y when
    aff is None (utils/misc.py:1051-1115)."""
    _need_cuda(X, "X")
    X4 = (X if X.dim() == 4 else X[..., None]).contiguous().float()
    out, factors = _zoom_to(X4, newsize)
    Y = out[..., 0] if out.shape[3] == 1 else out
    if aff is not None:
        return Y, _new_affine(aff, factors)
    return Y


def torch_resize(I, aff, resolution, power_factor_at_half_width=5, dtype=torch.float32, slow=False):
    """Resample a (X, Y, Z[, C]) volume with voxel-to-world matrix `aff` to `resolution` mm: per-axis Gaussian
    anti-aliasing blur (sigma = ln(power_factor) / pi * shape / newsize, none when not down-sampling; half width
    ceil(2.5 sigma); zero padding) followed by myzoom_torch_anisotropic (utils/misc.py:1117-1187).  `slow` is accepted
    for signature compatibility: there is one (GPU) path."""
    if dtype != torch.float32:
        raise NotImplementedError("torch_resize computes in float32")
    _need_cuda(I, "I")
    if I.dim() not in (3, 4):
        raise Exception('torch_resize works with 3D or 3D+label volumes')
    voxsize = np.sqrt(np.sum(aff[:-1, :-1] ** 2, axis=0))
    newsize = np.round(np.array(I.shape[0:3]) * (voxsize / resolution)).astype(int)
    factors = np.array(I.shape[0:3]) / np.array(newsize)
    k = np.log(power_factor_at_half_width) / np.pi
    sigmas = k * factors
    sigmas[sigmas <= k] = 0
    no_channels = I.dim() == 3
    vols = [I] if no_channels else list(I.unbind(3))
    L = _lib.lib()
    outs = []
    aff2 = None
    for V in vols:
        x = V.contiguous().to(torch.float32)
        nx, ny, nz = x.shape
        for d in range(3):
            if sigmas[d] > 0:
                sl = int(np.ceil(sigmas[d] * 2.5))
                v = np.arange(-sl, sl + 1)
                gauss = np.exp((-(v / sigmas[d]) ** 2 / 2))
                taps = torch.tensor(gauss / np.sum(gauss), device=x.device, dtype=torch.float32)
                y = torch.empty_like(x)
                _lib.check(L.bfm_blur_axis(x.data_ptr(), y.data_ptr(), nx, ny, nz, d, taps.data_ptr(), sl, _stream()))
                x = y
        out, f2 = _zoom_to(x[..., None], newsize)
        outs.append(out[..., 0])
        aff2 = _new_affine(aff, f2)
    res = outs[0] if no_channels else torch.stack(outs, dim=3)
    return res, aff2
