"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's spline resampling
(utils/interpol = vendored torch-interpol 0.2.3).  Never imported by the product package.

Covers the forward semantics of grid_pull / grid_push / grid_count / grid_grad for any dimension, spline orders
0..3, all seven boundary conditions and the three extrapolation modes, plus the recursive prefilter for orders
2..3 (utils/interpol/nd.py:31-288, bounds.py:30-89, splines.py:30-110, jit_utils.py:242-255, coeff.py:35-281).
Pinned against the reference itself by oracle/make_golden_interpol.py -> tests/golden/interpol.npz
(tests/test_interpol_oracle.py); higher orders are checked against those reference fixtures directly."""
import itertools
import math

import numpy as np

BOUNDS = {'zero': 0, 'replicate': 1, 'dct1': 2, 'dct2': 3, 'dst1': 4, 'dst2': 5, 'dft': 6}


def bound_index(t, i, n):
    """Bound.index (bounds.py:30-60) on int64 arrays."""
    i = i.astype(np.int64)
    if t in (0, 1):
        return np.clip(i, 0, n - 1)
    if t in (3, 5):
        n2 = 2 * n
        i = np.where(i < 0, n2 - 1 - np.mod(-i - 1, n2), np.mod(i, n2))
        return np.where(i >= n, n2 - 1 - i, i)
    if t == 2:
        if n == 1:
            return np.zeros_like(i)
        n2 = 2 * (n - 1)
        i = np.mod(np.abs(i), n2)
        return np.where(i >= n, n2 - i, i)
    if t == 4:
        n2 = 2 * (n + 1)
        i = np.where(i < 0, -i - 2, i)
        i = np.mod(i, n2)
        i = np.where(i > n, n2 - 2 - i, i)
        i = np.where(i == -1, 0, i)
        return np.where(i == n, n - 1, i)
    if t == 6:
        return np.mod(i, n)
    raise ValueError(t)


def bound_sign(t, i, n):
    """Bound.transform (bounds.py:62-89); 1 where the reference returns None."""
    i = i.astype(np.int64)
    if t == 4:
        if n == 1:
            return np.ones_like(i)
        n2 = 2 * (n + 1)
        i = np.where(i < 0, -i + (n - 1), i)
        i = np.mod(i, n2)
        x = np.where(i == 0, 0, 1)
        x = np.where(np.mod(i, n + 1) == n, 0, x)
        i = np.floor_divide(i, n + 1)
        return np.where(np.mod(i, 2) > 0, -x, x)
    if t == 5:
        i = np.where(i < 0, n - 1 - i, i)
        i = np.floor_divide(i, n)
        return np.where(np.mod(i, 2) > 0, -1, 1)
    if t == 0:
        return np.where((i < 0) | (i >= n), 0, 1)
    return np.ones_like(i)


def weight(o, x):
    """Spline.fastweight, orders 0..3 (splines.py:30-45)."""
    x = np.abs(x)
    if o == 0:
        return np.ones_like(x)
    if o == 1:
        return 1 - x
    if o == 2:
        return np.where(x < 0.5, 0.75 - x * x, 0.5 * (1.5 - x) ** 2)
    if o == 3:
        return np.where(x < 1., (x * x * (x - 2.) * 3. + 4.) / 6., (2. - x) ** 3 / 6.)
    raise NotImplementedError(o)


def dweight(o, x):
    """Spline.fastgrad, orders 0..3 (splines.py:90-105)."""
    s, x = np.sign(x), np.abs(x)
    if o == 0:
        return np.zeros_like(x)
    if o == 1:
        g = np.ones_like(x)
    elif o == 2:
        g = np.where(x < 0.5, -2 * x, x - 1.5)
    elif o == 3:
        g = np.where(x < 1, x * (x * 1.5 - 2), -0.5 * (2 - x) ** 2)
    else:
        raise NotImplementedError(o)
    return g * s


def _nodes(grid, shape, order, bound, all0):
    """Per-dimension node indices, weights (sign folded in) and derivative weights.
    When every order is 1 the reference dispatches to iso1 whose gradient is (+upper - lower) (iso1.py:269-387);
    the generic nd path uses sign(dist) for order-1 axes (splines.py:90-97) -- both are reproduced as they are."""
    all1 = all(o == 1 for o in order)
    idx, w, gw = [], [], []
    for d, (n, o, b) in enumerate(zip(shape, order, bound)):
        g = grid[..., d]
        if o == 0 and all0:
            f0 = np.rint(g)                                    # iso0.py:10-15 (torch.round)
        else:
            f0 = np.floor(g - (o - 1) / 2)                     # nd.py:45
        dist0 = g - f0
        i0 = f0.astype(np.int64)
        ii, ww, gg = [], [], []
        for k in range(o + 1):
            sg = bound_sign(b, i0 + k, n).astype(grid.dtype)
            ii.append(bound_index(b, i0 + k, n))
            ww.append((weight(o, dist0 - k) if o else np.ones_like(g)) * sg)
            gg.append((np.full_like(g, -1. if k == 0 else 1.) if all1 else dweight(o, dist0 - k)) * sg)
        idx.append(ii); w.append(ww); gw.append(gg)
    return idx, w, gw


def _mask(grid, shape, extrapolate):
    if extrapolate == 1:
        return np.ones(grid.shape[:-1], dtype=grid.dtype)
    thr = 0.55 if extrapolate == 2 else 0.05                  # jit_utils.py:242-255
    m = np.ones(grid.shape[:-1], dtype=bool)
    for d, n in enumerate(shape):
        m &= (grid[..., d] > -thr) & (grid[..., d] < n - 1 + thr)
    return m.astype(grid.dtype)


def pull(inp, grid, order, bound, extrapolate, grad=False):
    """inp (B,C,*ishape), grid (B,*oshape,D) -> (B,C,*oshape) or (B,C,*oshape,D)."""
    D = grid.shape[-1]
    B, Cn = inp.shape[:2]
    shape = inp.shape[2:]
    oshape = grid.shape[1:-1]
    g = grid.reshape(B, -1, D)
    idx, w, gw = _nodes(g, shape, order, bound, all(o == 0 for o in order))
    m = _mask(g, shape, extrapolate)
    flat = inp.reshape(B, Cn, -1)
    out = np.zeros((B, Cn, g.shape[1]) + ((D,) if grad else ()), dtype=inp.dtype)
    strides = [int(np.prod(shape[d + 1:])) for d in range(D)]
    for nodes in itertools.product(*[range(o + 1) for o in order]):
        lin = sum(idx[d][k] * strides[d] for d, k in enumerate(nodes))
        v = np.take_along_axis(flat, np.broadcast_to(lin[:, None, :], (B, Cn, lin.shape[-1])), axis=2)
        if not grad:
            ww = np.ones_like(g[..., 0])
            for d, k in enumerate(nodes):
                ww = ww * w[d][k]
            out += v * ww[:, None, :]
        else:
            for dd in range(D):
                ww = np.ones_like(g[..., 0])
                for d, k in enumerate(nodes):
                    ww = ww * (gw[d][k] if d == dd else w[d][k])
                out[..., dd] += v * ww[:, None, :]
    out = out * (m[:, None, :, None] if grad else m[:, None, :])
    return out.reshape((B, Cn) + tuple(oshape) + ((D,) if grad else ()))


def push(inp, grid, shape, order, bound, extrapolate):
    """inp (B,C,*ishape) or None (count), grid (B,*ishape,D) -> (B,C,*shape)."""
    D = grid.shape[-1]
    B = grid.shape[0]
    g = grid.reshape(B, -1, D)
    Cn = 1 if inp is None else inp.shape[1]
    vals = np.ones((B, 1, g.shape[1]), dtype=grid.dtype) if inp is None else inp.reshape(B, Cn, -1)
    idx, w, _ = _nodes(g, shape, order, bound, all(o == 0 for o in order))
    m = _mask(g, shape, extrapolate)
    out = np.zeros((B, Cn, int(np.prod(shape))), dtype=vals.dtype)
    strides = [int(np.prod(shape[d + 1:])) for d in range(D)]
    for nodes in itertools.product(*[range(o + 1) for o in order]):
        lin = sum(idx[d][k] * strides[d] for d, k in enumerate(nodes))
        ww = m.copy()
        for d, k in enumerate(nodes):
            ww = ww * w[d][k]
        for b in range(B):
            for c in range(Cn):
                np.add.at(out[b, c], lin[b], vals[b, c] * ww[b])
    return out.reshape((B, Cn) + tuple(shape))


POLES = {2: [math.sqrt(8.) - 3.], 3: [math.sqrt(3.) - 2.]}


def spline_filter(x, order, bound, axis):
    """coeff.filter (coeff.py:255-281) along one axis; bound in {0,1,2,3,6}."""
    if order < 2 or x.shape[axis] == 1:
        return x.copy()
    c = np.moveaxis(x.copy(), axis, 0)
    n = c.shape[0]
    T = c.dtype.type
    poles = POLES[order]
    gain = 1.
    for p in poles:
        gain *= (1. - p) * (1. - 1. / p)
    c *= T(gain)
    for pole in poles:
        max_iter = int(math.ceil(-30. / math.log(abs(pole))))
        if bound in (0, 2):
            if max_iter < n:
                pw = (T(pole) ** np.arange(1, max_iter)).astype(c.dtype)
                init = np.tensordot(pw, c[1:max_iter], axes=(0, 0)) + c[0]
            else:
                polen = pole ** (n - 1)
                pw = (T(pole) ** np.arange(1, n - 1)).astype(c.dtype)
                pw = pw + T(polen * polen) / pw
                init = np.tensordot(pw, c[1:-1], axes=(0, 0)) + (c[0] + T(polen) * c[-1])
                init = init / T(1 - polen * polen)
        elif bound in (1, 3):
            polen = pole ** n
            pole_last = polen * (1 + 1 / (pole + polen * polen))
            pw = ((T(pole) ** np.arange(1, n - 1)) + (T(pole) ** np.arange(2 * n - 2, n, -1))).astype(c.dtype)
            init = np.tensordot(pw, c[1:-1], axes=(0, 0)) + (c[0] + T(pole_last) * c[-1])
            init = init * T(pole / (1 - polen * polen)) + c[0]
        else:
            mi = min(max_iter, n)
            pw = (T(pole) ** np.arange(1, mi)).astype(c.dtype)[::-1]
            init = (np.tensordot(pw, c[n - mi + 1:], axes=(0, 0)) + c[0]) / T(1 - pole ** mi)
        c[0] = init
        for i in range(1, n):
            c[i] = c[i] + T(pole) * c[i - 1]
        if bound in (0, 2):
            fin = (T(pole) * c[-2] + c[-1]) * T(pole / (pole * pole - 1))
        elif bound in (1, 3):
            fin = c[-1] * T(pole / (pole - 1))
        else:
            mi = min(max_iter, n)
            pw = (T(pole) ** np.arange(2, mi + 1)).astype(c.dtype)
            fin = (np.tensordot(pw, c[:mi - 1], axes=(0, 0)) + T(pole) * c[-1]) / T(pole ** mi - 1)
        c[-1] = fin
        for i in range(n - 2, -1, -1):
            c[i] = (c[i + 1] - c[i]) * T(pole)
    return np.moveaxis(c, 0, axis)
