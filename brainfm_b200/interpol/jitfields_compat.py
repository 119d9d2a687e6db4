"""A `jitfields`-shaped front end of libbfm's spline kernels, so that the REFERENCE's own `utils.interpol` can
forward to them through the hook it already has.

The reference dispatches every public entry point to the third-party `jitfields` package when
`utils/interpol/backend.py:1` (`jitfields = False`) is switched on and `import jitfields` succeeds
(utils/interpol/jitfields.py:1-7, api.py:174, 236, 282, 323, 378, 428; resize.py:66; restrict.py:62).  This module
provides the functions that shim calls (utils/interpol/jitfields.py:28-95), with jitfields' conventions:

    channels LAST:  pull / push / grad take (..., *spatial, channel) tensors and (..., *spatial', ndim) grids
    order= / ndim=  instead of interpolation= / dim=

To install it for a reference checkout:

    import sys, brainfm_b200.interpol.jitfields_compat as jf
    sys.modules['jitfields'] = jf            # before `import utils.interpol`
    import utils.interpol as interpol
    interpol.backend.jitfields = True        # every grid_pull / grid_push / ... now runs in libbfm on the GPU

Tensors must live on a CUDA device (there is no CPU path).
"""
import torch

from . import api

__all__ = ['pull', 'push', 'count', 'grad', 'spline_coeff', 'spline_coeff_', 'spline_coeff_nd', 'spline_coeff_nd_',
           'resize', 'restrict']


def _first(x, ndim):
    """(..., *spatial, channel) -> (..., channel, *spatial)"""
    return torch.movedim(x, -1, -ndim - 1)


def _last(x, ndim, grad=False):
    """(..., channel, *spatial[, ndim]) -> (..., *spatial, channel[, ndim])"""
    return torch.movedim(x, -ndim - 1 - grad, -1 - grad)


def _finish(res, out):
    if out is not None:
        out.copy_(res)
        return out
    return res


def pull(inp, grid, order=2, bound='dct2', extrapolate=True, prefilter=False, out=None):
    """inp (..., *inshape, channel), grid (..., *outshape, ndim) -> (..., *outshape, channel)"""
    ndim = grid.shape[-1]
    res = api.grid_pull(_first(inp, ndim), grid, interpolation=order, bound=bound, extrapolate=extrapolate,
                        prefilter=prefilter)
    return _finish(_last(res, ndim), out)


def push(inp, grid, shape=None, order=2, bound='dct2', extrapolate=True, prefilter=False, out=None):
    """inp (..., *inshape, channel), grid (..., *inshape, ndim) -> (..., *shape, channel)"""
    ndim = grid.shape[-1]
    res = api.grid_push(_first(inp, ndim), grid, shape, interpolation=order, bound=bound, extrapolate=extrapolate,
                        prefilter=prefilter)
    return _finish(_last(res, ndim), out)


def count(grid, shape=None, order=2, bound='dct2', extrapolate=True, out=None):
    """grid (..., *inshape, ndim) -> (..., *shape)"""
    return _finish(api.grid_count(grid, shape, interpolation=order, bound=bound, extrapolate=extrapolate), out)


def grad(inp, grid, order=2, bound='dct2', extrapolate=True, prefilter=False, out=None):
    """inp (..., *inshape, channel), grid (..., *outshape, ndim) -> (..., *outshape, channel, ndim)"""
    ndim = grid.shape[-1]
    res = api.grid_grad(_first(inp, ndim), grid, interpolation=order, bound=bound, extrapolate=extrapolate,
                        prefilter=prefilter)
    return _finish(_last(res, ndim, True), out)


def spline_coeff(inp, order, bound='dct2', dim=-1):
    return api.spline_coeff(inp, interpolation=order, bound=bound, dim=dim, inplace=False)


def spline_coeff_(inp, order, bound='dct2', dim=-1):
    return api.spline_coeff(inp, interpolation=order, bound=bound, dim=dim, inplace=True)


def spline_coeff_nd(inp, order, bound='dct2', ndim=None):
    return api.spline_coeff_nd(inp, interpolation=order, bound=bound, dim=ndim, inplace=False)


def spline_coeff_nd_(inp, order, bound='dct2', ndim=None):
    return api.spline_coeff_nd(inp, interpolation=order, bound=bound, dim=ndim, inplace=True)


def resize(x, factor=None, shape=None, ndim=None, anchor='e', order=2, bound='dct2', prefilter=True):
    """x (batch, channel, *spatial): the reference's shim passes its channel-first image through unchanged
    (utils/interpol/jitfields.py:76-84)."""
    return api.resize(x, factor=factor, shape=shape, anchor=anchor, interpolation=order, prefilter=prefilter,
                      bound=bound)


def restrict(x, factor=None, shape=None, ndim=None, anchor='e', order=1, bound='dct2', reduce_sum=False):
    return api.restrict(x, factor=factor, shape=shape, anchor=anchor, interpolation=order, reduce_sum=reduce_sum,
                        bound=bound)
