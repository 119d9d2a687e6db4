"""Finite differences and the curl ("stream") velocity (mirror of ShapeID/misc.py:66-80, 84-259)."""
import torch

from .. import _lib
from ._common import fvec, ivec, need_cuda, stream


def _gradient(X, mode, batched, delta_lst):
    need_cuda(X, "X")
    dim = X.dim() - 1 if batched else X.dim()
    if dim != 3:
        raise NotImplementedError("brainfm_b200.ShapeID gradients are implemented for 3-D volumes")
    if X.dtype not in (torch.float32, torch.float64):
        X = X.float()
    Xs = X if batched else X[None]
    out = torch.empty((*Xs.shape, 3), dtype=torch.float32, device=X.device)
    L = _lib.lib()
    for b in range(Xs.shape[0]):
        xb = Xs[b].contiguous()
        _lib.check(L.bfm_gradient3d(xb.data_ptr(), 1 if xb.dtype == torch.float64 else 0, ivec(xb.shape), mode,
                                    fvec(delta_lst), out[b].data_ptr(), stream()))
    return out if batched else out[0]


def gradient_c(X, batched=False, delta_lst=[1., 1., 1.]):
    """Central differences, one-sided at the borders, float32 output (ShapeID/misc.py:198-259)."""
    return _gradient(X, 0, batched, delta_lst)


def gradient_f(X, batched=False, delta_lst=[1., 1., 1.]):
    """Forward differences, backward at the upper border (ShapeID/misc.py:84-139, DiffEqs/pde.py:13-67)."""
    return _gradient(X, 1, batched, delta_lst)


def gradient_b(X, batched=False, delta_lst=[1., 1., 1.]):
    """Backward differences, forward at the lower border (ShapeID/misc.py:141-196, DiffEqs/pde.py:70-124)."""
    return _gradient(X, 2, batched, delta_lst)


def stream_3D(Phi_a, Phi_b, Phi_c, batched=False, delta_lst=[1., 1., 1.], multiplier=1.0):
    """Divergence-free velocity = curl of three potentials (ShapeID/misc.py:66-80), one fused kernel."""
    for t in (Phi_a, Phi_b, Phi_c):
        need_cuda(t, "Phi")
    if batched or list(delta_lst) != [1., 1., 1.]:
        dDa = gradient_c(Phi_a, batched, delta_lst)
        dDb = gradient_c(Phi_b, batched, delta_lst)
        dDc = gradient_c(Phi_c, batched, delta_lst)
        return ((dDc[..., 1] - dDb[..., 2]) * multiplier, (dDa[..., 2] - dDc[..., 0]) * multiplier,
                (dDb[..., 0] - dDa[..., 1]) * multiplier)
    dt = Phi_a.dtype if Phi_a.dtype in (torch.float32, torch.float64) else torch.float32
    a, b, c = (t.to(dt).contiguous() for t in (Phi_a, Phi_b, Phi_c))
    V = torch.empty((3, *a.shape), dtype=torch.float32, device=a.device)
    _lib.check(_lib.lib().bfm_curl3d(a.data_ptr(), b.data_ptr(), c.data_ptr(), 1 if dt == torch.float64 else 0,
                                     ivec(a.shape), float(multiplier), V[0].data_ptr(), V[1].data_ptr(),
                                     V[2].data_ptr(), stream()))
    return V[0], V[1], V[2]


def center_crop(img, win_size=[220, 220, 220]):
    """Centre crop / no-op when the window is larger (ShapeID/misc.py:10-38)."""
    shp = img.shape[-3:]
    st = [max(0, (s - w) // 2) for s, w in zip(shp, win_size)]
    return img[..., st[0]:st[0] + win_size[0], st[1]:st[1] + win_size[1], st[2]:st[2] + win_size[2]]
