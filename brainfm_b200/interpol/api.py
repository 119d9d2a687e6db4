"""High-level resampling API (mirror of utils/interpol/api.py, resize.py, restrict.py) over libbfm.

float32 and float64; 1-D, 2-D and 3-D (lower dimensions run through the 3-D kernel with singleton axes).
Differentiable like the reference (utils/interpol/autograd.py:125-301): grid_pull / grid_push / grid_count /
spline_coeff(_nd) carry autograd Functions whose backward passes are the adjoint kernels (pull <-> push, grid
gradients through the grid_grad kernel); grid_grad itself is forward-only.  Input layout as in the reference:
input (..., [channel], *spatial) channels-first, grid (..., *spatial_out, dim) in voxel coordinates."""
import ctypes as C
import math

import torch

from .. import _lib

_BOUNDS = {'zero': 0, 'zeros': 0, 'constant': 0, 'replicate': 1, 'repeat': 1, 'border': 1, 'nearest': 1,
           'dct1': 2, 'mirror': 2, 'dct2': 3, 'reflect': 3, 'reflection': 3, 'neumann': 3,
           'dst1': 4, 'antimirror': 4, 'dst2': 5, 'antireflect': 5, 'dirichlet': 5, 'dft': 6, 'wrap': 6,
           'circular': 6}
_ORDERS = {'nearest': 0, 'linear': 1, 'quadratic': 2, 'cubic': 3, 'fourth': 4, 'fifth': 5, 'sixth': 6,
           'seventh': 7}


def _stream():
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


def _as_list(x, n):
    x = list(x) if isinstance(x, (list, tuple)) else [x]
    return (x + [x[-1]] * n)[:n]          # pad_list_int (jit_utils.py:10-15)


def _bounds(bound, dim):
    out = []
    for b in _as_list(bound, dim):
        if hasattr(b, 'value'):
            b = b.value
        if isinstance(b, str):
            if b.lower() not in _BOUNDS:
                raise ValueError(f'Unknown boundary condition {b}')
            out.append(_BOUNDS[b.lower()])
        elif isinstance(b, int) and 0 <= b <= 6:
            out.append(b)
        else:
            raise ValueError(f'Unknown boundary condition {b}')
    return out


def _orders(inter, dim):
    out = []
    for o in _as_list(inter, dim):
        if hasattr(o, 'value'):
            o = o.value
        if isinstance(o, str):
            if o.lower() not in _ORDERS:
                raise ValueError(f'Unknown interpolation order {o}')
            out.append(_ORDERS[o.lower()])
        elif isinstance(o, int) and 0 <= o <= 7:
            out.append(o)
        else:
            raise ValueError(f'Unknown interpolation order {o}')
    return out


def _extrap(e):
    if isinstance(e, bool):
        return 1 if e else 0
    if isinstance(e, str):
        return {'no': 0, 'yes': 1, 'hist': 2}[e]
    return int(e)


def _expanded_shape(*shapes):
    n = max(len(s) for s in shapes)
    out = [1] * n
    for s in shapes:
        s = [1] * (n - len(s)) + list(s)
        for i, v in enumerate(s):
            if v != 1:
                if out[i] != 1 and out[i] != v:
                    raise ValueError('Incompatible shapes for broadcasting')
                out[i] = v
    return out


def _need(t, name):
    if not t.is_cuda:
        raise _lib.BfmError("%s must be a CUDA tensor: brainfm_b200.interpol has no CPU path" % name)
    if t.dtype not in (torch.float32, torch.float64):
        raise NotImplementedError("brainfm_b200.interpol supports float32 and float64 (got %s)" % t.dtype)


def _pad3(shape):
    return [1] * (3 - len(shape)) + list(shape)


def _run(mode, inp, grid, out, ishape, order, bound, extrapolate, B, Cn, Bi, Bg, P):
    dim = len(ishape)
    pad = 3 - dim
    ish = (C.c_int * 3)(*_pad3(ishape))
    od = (C.c_int * 3)(*([0] * pad + order))
    bd = (C.c_int * 3)(*([1] * pad + bound))
    iso = 1 if all(o == 0 for o in order) else (2 if all(o == 1 for o in order) else 0)   # pushpull.py:35-60
    _lib.check(_lib.lib().bfm_interpol(mode, 1 if grid.dtype == torch.float64 else 0,
                                       None if inp is None else inp.data_ptr(), grid.data_ptr(), out.data_ptr(),
                                       ish, od, bd, extrapolate, iso, B, Cn, Bi, Bg, P, _stream()))


def _grid3(grid):
    """(B, P, dim) contiguous grid -> (B, P, 3) with zero coordinates on the padded leading axes."""
    dim = grid.shape[-1]
    if dim == 3:
        return grid.contiguous()
    g = grid.new_zeros([*grid.shape[:-1], 3])
    g[..., 3 - dim:] = grid
    return g


def _preproc(grid, input=None, mode=None):
    dim = grid.shape[-1]
    if input is None:
        spatial = list(grid.shape[-dim - 1:-1])
        batch = list(grid.shape[:-dim - 1])
        grid = grid.reshape([-1, *spatial, dim])
        return grid, dict(batch=batch, channel=[1] if batch else [], dim=dim)
    grid_spatial = list(grid.shape[-dim - 1:-1])
    grid_batch = list(grid.shape[:-dim - 1])
    input_spatial = list(input.shape[-dim:])
    channel = 0 if input.dim() == dim else input.shape[-dim - 1]
    input_batch = list(input.shape[:-dim - 1])
    if mode == 'push':
        grid_spatial = input_spatial = _expanded_shape(grid_spatial, input_spatial)
    batch = _expanded_shape(grid_batch, input_batch)
    grid = grid.expand([*batch, *grid_spatial, dim]).reshape([-1, *grid_spatial, dim])
    input = input.expand([*batch, channel or 1, *input_spatial]).reshape([-1, channel or 1, *input_spatial])
    out_channel = [channel] if channel else ([1] if batch else [])
    return grid, input, dict(batch=batch, channel=out_channel, dim=dim)


def _postproc(out, info, mode):
    dim = info['dim']
    if mode != 'grad':
        spatial, feat = list(out.shape[-dim:]), []
    else:
        spatial, feat = list(out.shape[-dim - 1:-1]), [out.shape[-1]]
    return out.reshape([*info['batch'], *info['channel'], *spatial, *feat])


def _pull_raw(input, grid, order, bound, extrapolate, mode=0):
    """input (B, C, *ishape), grid (B, *oshape, dim) -> (B, C, *oshape[, dim])"""
    dim = grid.shape[-1]
    B, Cn = input.shape[:2]
    ishape = list(input.shape[2:])
    oshape = list(grid.shape[1:-1])
    P = int(math.prod(oshape))
    dt = torch.promote_types(input.dtype, grid.dtype)
    if (mode == 0 and dim == 3 and dt == torch.float32 and order[0] in (1, 3) and order[0] == order[1] == order[2]
            and all(s >= 0 for s in input.stride())
            and sum((n - 1) * s for n, s in zip(input.shape[1:], input.stride()[1:])) < 2 ** 31):
        # fast path: one compile-time order, the input read in place through its strides (channels-last views too)
        inp = input.to(dt)
        # the memory format of the result follows the input: a channels-last view in, a channels-last view out
        chlast = Cn > 1 and inp.stride(1) == 1
        if order[0] == 3 and Cn in (2, 3, 4) and not chlast:
            # 64 taps per point: one transposition to channels-last turns C scalar gathers per tap into one vector load
            inp = inp.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)
        g = grid.to(dt).reshape(B, P, dim).contiguous()
        if chlast:
            out = torch.empty((B, *oshape, Cn), dtype=dt, device=inp.device).permute(0, 4, 1, 2, 3)
        else:
            out = torch.empty((B, Cn, *oshape), dtype=dt, device=inp.device)
        _lib.check(_lib.lib().bfm_interpol_pull_fast(
            inp.data_ptr(), (C.c_int64 * 5)(*inp.stride()), g.data_ptr(), P * 3, out.data_ptr(), int(chlast),
            (C.c_int * 3)(*ishape), order[0], (C.c_int * 3)(*bound), extrapolate, B, Cn, P, _stream()))
        return out
    inp = input.to(dt).contiguous()
    g = _grid3(grid.to(dt).reshape(B, P, dim))
    if mode == 0:
        out = torch.empty((B, Cn, *oshape), dtype=dt, device=inp.device)
    else:
        out3 = torch.empty((B, Cn, P, 3), dtype=dt, device=inp.device)
        out = out3
    _run(mode, inp, g, out, ishape, order, bound, extrapolate, B, Cn, B, B, P)
    if mode == 2:
        out = out3[..., 3 - dim:].reshape(B, Cn, *oshape, dim)
    return out


def _wants_grad(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


class _Pull(torch.autograd.Function):
    """grid_pull with its adjoints (utils/interpol/autograd.py:125-155, pushpull.py grid_pull_backward):
    d/d input = push of the incoming gradient, d/d grid = sum_c grad_c * (spatial gradient of input_c at grid)."""

    @staticmethod
    def forward(ctx, input, grid, order, bound, ext):
        ctx.opt = (order, bound, ext)
        ctx.save_for_backward(input, grid)
        return _pull_raw(input, grid, order, bound, ext)

    @staticmethod
    def backward(ctx, grad):
        input, grid = ctx.saved_tensors
        order, bound, ext = ctx.opt
        gi = gg = None
        grad = grad.contiguous()
        if ctx.needs_input_grad[0]:
            gi = _push_raw(grad, grid, list(input.shape[2:]), order, bound, ext).to(input.dtype)
        if ctx.needs_input_grad[1]:
            g3 = _pull_raw(input, grid, order, bound, ext, mode=2)             # (B, C, *out, dim)
            gg = (g3 * grad.unsqueeze(-1)).sum(1).to(grid.dtype)
        return gi, gg, None, None, None


class _Grad(torch.autograd.Function):
    """grid_grad with its adjoints (utils/interpol/autograd.py:216-243, pushpull.py grid_grad_backward):
    d/d input = push of the incoming gradient with the derivative weights (grid_pushgrad), d/d grid = the spline
    Hessian of the input contracted with the incoming gradient (grid_hess)."""

    @staticmethod
    def forward(ctx, input, grid, order, bound, ext):
        ctx.opt = (order, bound, ext)
        ctx.save_for_backward(input, grid)
        return _pull_raw(input, grid, order, bound, ext, mode=2)

    @staticmethod
    def backward(ctx, grad):
        input, grid = ctx.saved_tensors
        order, bound, ext = ctx.opt
        dim = grid.shape[-1]
        B, Cn = input.shape[:2]
        ishape = list(input.shape[2:])
        oshape = list(grid.shape[1:-1])
        P = int(math.prod(oshape))
        dt = torch.promote_types(input.dtype, grid.dtype)
        inp = input.to(dt).contiguous()
        g = _grid3(grid.to(dt).reshape(B, P, dim))
        go = grad.to(dt).reshape(B, Cn, P, dim)
        if dim < 3:                                   # padded leading axes: zero incoming gradient there
            g3 = go.new_zeros(B, Cn, P, 3)
            g3[..., 3 - dim:] = go
            go = g3
        go = go.contiguous()
        gi = torch.zeros_like(inp) if ctx.needs_input_grad[0] else None
        gg = torch.empty((B, P, 3), dtype=dt, device=inp.device) if ctx.needs_input_grad[1] else None
        pad = 3 - dim
        iso = 1 if all(o == 0 for o in order) else (2 if all(o == 1 for o in order) else 0)
        _lib.check(_lib.lib().bfm_interpol_grad_backward(
            1 if dt == torch.float64 else 0, go.data_ptr(), inp.data_ptr(), g.data_ptr(),
            None if gi is None else gi.data_ptr(), None if gg is None else gg.data_ptr(),
            (C.c_int * 3)(*_pad3(ishape)), (C.c_int * 3)(*([0] * pad + order)), (C.c_int * 3)(*([1] * pad + bound)),
            ext, iso, B, Cn, P, _stream()))
        if gi is not None:
            gi = gi.to(input.dtype)
        if gg is not None:
            gg = gg[..., 3 - dim:].reshape(B, *oshape, dim).to(grid.dtype)
        return gi, gg, None, None, None


class _Push(torch.autograd.Function):
    """grid_push (autograd.py:158-188): d/d input = pull of the incoming gradient, d/d grid = sum_c input_c *
    (spatial gradient of grad_c at grid)."""

    @staticmethod
    def forward(ctx, input, grid, shape, order, bound, ext):
        ctx.opt = (order, bound, ext)
        ctx.save_for_backward(input, grid)
        return _push_raw(input, grid, list(shape), order, bound, ext)

    @staticmethod
    def backward(ctx, grad):
        input, grid = ctx.saved_tensors
        order, bound, ext = ctx.opt
        gi = gg = None
        grad = grad.contiguous()
        if ctx.needs_input_grad[0]:
            gi = _pull_raw(grad, grid, order, bound, ext).to(input.dtype)
        if ctx.needs_input_grad[1]:
            g3 = _pull_raw(grad, grid, order, bound, ext, mode=2)
            gg = (g3 * input.unsqueeze(-1)).sum(1).to(grid.dtype)
        return gi, gg, None, None, None, None


class _Count(torch.autograd.Function):
    """grid_count = push of ones (autograd.py:191-219)."""

    @staticmethod
    def forward(ctx, grid, shape, order, bound, ext):
        ctx.opt = (order, bound, ext)
        ctx.save_for_backward(grid)
        return _push_raw(None, grid, list(shape), order, bound, ext)

    @staticmethod
    def backward(ctx, grad):
        grid, = ctx.saved_tensors
        order, bound, ext = ctx.opt
        gg = None
        if ctx.needs_input_grad[0]:
            gg = _pull_raw(grad.contiguous(), grid, order, bound, ext, mode=2).sum(1).to(grid.dtype)
        return gg, None, None, None, None


class _Coeff(torch.autograd.Function):
    """spline_coeff / spline_coeff_nd: the prefilter is symmetric, backward == forward (autograd.py:254-301)."""

    @staticmethod
    def forward(ctx, input, fn, args):
        ctx.fn, ctx.args = fn, args
        return fn(input, *args, inplace=False)

    @staticmethod
    def backward(ctx, grad):
        return ctx.fn(grad.contiguous(), *ctx.args, inplace=False), None, None


def grid_pull(input, grid, interpolation='linear', bound='zero', extrapolate=False, prefilter=False):
    """Sample an image with respect to a deformation field (utils/interpol/api.py:137-200)."""
    _need(grid, 'grid')
    dim = grid.shape[-1]
    order, bnd, ext = _orders(interpolation, dim), _bounds(bound, dim), _extrap(extrapolate)
    grid, input, info = _preproc(grid, input)
    if not input.dtype.is_floating_point:
        # label map: soft-pull every label, keep the arg-max (api.py:182-193)
        out = input.new_zeros([*input.shape[:2], *grid.shape[1:-1]])
        pmax = grid.new_zeros([*input.shape[:2], *grid.shape[1:-1]])
        for label in input.unique():
            soft = (input == label).to(grid.dtype)
            if prefilter:
                soft = spline_coeff_nd(soft, interpolation=interpolation, bound=bound, dim=dim, inplace=True)
            soft = _pull_raw(soft, grid, order, bnd, ext)
            out[soft > pmax] = label
            pmax = torch.max(pmax, soft)
    else:
        _need(input, 'input')
        if prefilter:
            input = spline_coeff_nd(input, interpolation=interpolation, bound=bound, dim=dim)
        if _wants_grad(input, grid):
            out = _Pull.apply(input, grid, order, bnd, ext)
        else:
            out = _pull_raw(input, grid, order, bnd, ext)
    return _postproc(out, info, 'pull')


def grid_grad(input, grid, interpolation='linear', bound='zero', extrapolate=False, prefilter=False):
    """Sample spatial gradients of an image (utils/interpol/api.py:290-332)."""
    _need(grid, 'grid')
    _need(input, 'input')
    dim = grid.shape[-1]
    order, bnd, ext = _orders(interpolation, dim), _bounds(bound, dim), _extrap(extrapolate)
    grid, input, info = _preproc(grid, input)
    if prefilter:
        input = spline_coeff_nd(input, interpolation=interpolation, bound=bound, dim=dim)
    if _wants_grad(input, grid):
        out = _Grad.apply(input, grid, order, bnd, ext)
    else:
        out = _pull_raw(input, grid, order, bnd, ext, mode=2)
    return _postproc(out, info, 'grad')


def _push_raw(input, grid, shape, order, bound, extrapolate):
    dim = grid.shape[-1]
    B = grid.shape[0]
    P = int(math.prod(grid.shape[1:-1]))
    dt = grid.dtype if input is None else torch.promote_types(input.dtype, grid.dtype)
    g = _grid3(grid.to(dt).reshape(B, P, dim))
    Cn = 1 if input is None else input.shape[1]
    out = torch.zeros((B, Cn, *shape), dtype=dt, device=grid.device)
    inp = None if input is None else input.to(dt).reshape(B, Cn, P).contiguous()
    _run(1, inp, g, out, list(shape), order, bound, extrapolate, B, Cn, B, B, P)
    return out


def grid_push(input, grid, shape=None, interpolation='linear', bound='zero', extrapolate=False, prefilter=False):
    """Splat an image with respect to a deformation field (utils/interpol/api.py:203-250)."""
    _need(grid, 'grid')
    _need(input, 'input')
    dim = grid.shape[-1]
    order, bnd, ext = _orders(interpolation, dim), _bounds(bound, dim), _extrap(extrapolate)
    grid, input, info = _preproc(grid, input, mode='push')
    if shape is None:
        shape = tuple(input.shape[2:])
    if list(input.shape[2:]) != list(grid.shape[1:-1]):
        raise ValueError('Input and grid should have the same spatial shape')
    if _wants_grad(input, grid):
        out = _Push.apply(input, grid, tuple(shape), order, bnd, ext)
        if prefilter:
            out = spline_coeff_nd(out, interpolation=interpolation, bound=bound, dim=dim)
    else:
        out = _push_raw(input, grid, list(shape), order, bnd, ext)
        if prefilter:
            out = spline_coeff_nd(out, interpolation=interpolation, bound=bound, dim=dim, inplace=True)
    return _postproc(out, info, 'push')


def grid_count(grid, shape=None, interpolation='linear', bound='zero', extrapolate=False):
    """Splatting weights with respect to a deformation field (utils/interpol/api.py:253-287)."""
    _need(grid, 'grid')
    dim = grid.shape[-1]
    order, bnd, ext = _orders(interpolation, dim), _bounds(bound, dim), _extrap(extrapolate)
    grid, info = _preproc(grid)
    if shape is None:
        shape = tuple(grid.shape[1:-1])
    if _wants_grad(grid):
        out = _Count.apply(grid, tuple(shape), order, bnd, ext)
    else:
        out = _push_raw(None, grid, list(shape), order, bnd, ext)
    return _postproc(out, info, 'count')


_POLES = {
    2: [math.sqrt(8.) - 3.],
    3: [math.sqrt(3.) - 2.],
    4: [math.sqrt(664. - math.sqrt(438976.)) + math.sqrt(304.) - 19.,
        math.sqrt(664. + math.sqrt(438976.)) - math.sqrt(304.) - 19.],
    5: [math.sqrt(67.5 - math.sqrt(4436.25)) + math.sqrt(26.25) - 6.5,
        math.sqrt(67.5 + math.sqrt(4436.25)) - math.sqrt(26.25) - 6.5],
    6: [-0.488294589303044755130118038883789062112279161239377608394,
        -0.081679271076237512597937765737059080653379610398148178525368,
        -0.00141415180832581775108724397655859252786416905534669851652709],
    7: [-0.5352804307964381655424037816816460718339231523426924148812,
        -0.122554615192326690515272264359357343605486549427295558490763,
        -0.0091486948096082769285930216516478534156925639545994482648003],
}


def spline_coeff(input, interpolation='linear', bound='dct2', dim=-1, inplace=False):
    """Interpolating spline coefficients along one dimension (utils/interpol/api.py:335-383, coeff.py:255-316)."""
    _need(input, 'input')
    if _wants_grad(input):
        return _Coeff.apply(input, _spline_coeff_nograd, (interpolation, bound, dim))
    return _spline_coeff_nograd(input, interpolation, bound, dim, inplace)


def _spline_coeff_nograd(input, interpolation='linear', bound='dct2', dim=-1, inplace=False):
    order = _orders(interpolation, 1)[0]
    bnd = _bounds(bound, 1)[0]
    if input.requires_grad:
        input = input.detach()
    out = input if (inplace and input.is_contiguous()) else input.clone(memory_format=torch.contiguous_format)
    if order in (0, 1) or out.shape[dim] == 1:
        return out
    if bnd not in (0, 1, 2, 3, 6):
        raise NotImplementedError
    d = dim % out.dim()
    outer = int(math.prod(out.shape[:d]))
    inner = int(math.prod(out.shape[d + 1:]))
    poles = (C.c_double * 3)(*(_POLES[order] + [0.0] * (3 - len(_POLES[order]))))
    _lib.check(_lib.lib().bfm_spline_filter(out.data_ptr(), 1 if out.dtype == torch.float64 else 0, outer,
                                            int(out.shape[d]), inner, bnd, poles, len(_POLES[order]), _stream()))
    if inplace and out is not input:
        input.copy_(out)
        return input
    return out


def spline_coeff_nd(input, interpolation='linear', bound='dct2', dim=None, inplace=False):
    """Interpolating spline coefficients along the last `dim` dimensions (api.py:386-447, coeff.py:319-344)."""
    _need(input, 'input')
    if dim is None:
        dim = input.dim()
    if _wants_grad(input):
        return _Coeff.apply(input, _spline_coeff_nd_nograd, (interpolation, bound, dim))
    return _spline_coeff_nd_nograd(input, interpolation, bound, dim, inplace)


def _spline_coeff_nd_nograd(input, interpolation='linear', bound='dct2', dim=None, inplace=False):
    if input.requires_grad:
        input = input.detach()
    orders, bnds = _orders(interpolation, dim), _bounds(bound, dim)
    out = input if inplace else input.clone(memory_format=torch.contiguous_format)
    for d, (b, o) in enumerate(zip(bnds, orders)):
        out = _spline_coeff_nograd(out, o, b, dim=-dim + d, inplace=True)
    return out


def identity_grid(shape, dtype=None, device=None):
    """Identity deformation field (api.py:455-477)."""
    mesh1d = [torch.arange(float(s), dtype=dtype, device=device) for s in shape]
    return torch.stack(torch.meshgrid(*mesh1d, indexing='ij'), dim=-1)


def add_identity_grid_(disp):
    """Adds the identity grid to a displacement field, in place (api.py:480-504)."""
    dim = disp.shape[-1]
    spatial = disp.shape[-dim - 1:-1]
    mesh1d = [torch.arange(s, dtype=disp.dtype, device=disp.device) for s in spatial]
    for i, g in enumerate(torch.meshgrid(*mesh1d, indexing='ij')):
        disp[..., i].add_(g)
    return disp


def add_identity_grid(disp):
    """Adds the identity grid to a displacement field (api.py:507-521)."""
    if disp.is_cuda and disp.dtype == torch.float32 and disp.shape[-1] == 3 and disp.dim() >= 4 and disp.is_contiguous():
        out = torch.empty_like(disp)
        X, Y, Z = disp.shape[-4:-1]
        Bn = int(math.prod(disp.shape[:-4]))
        if Bn > 0 and disp.numel() > 0:
            _lib.check(_lib.lib().bfm_add_identity_grid(disp.data_ptr(), out.data_ptr(), Bn, X, Y, Z, _stream()))
        return out
    return add_identity_grid_(disp.clone())


def compose_step(disp, bound='dct2', extrapolate=True):
    """disp + grid_pull(disp, add_identity_grid(disp)) with linear interpolation: one squaring step of a displacement
    field (B..., X, Y, Z, 3), fused into one kernel (bfm_compose_step) -- bit-identical to the three-call form, one
    read and one write of the field instead of eight.  Extension (the reference's package composes it from its
    public calls, e.g. BASELINE configs[2]); other dtypes / dimensions fall back to exactly that composition."""
    if (disp.is_cuda and disp.dtype == torch.float32 and disp.shape[-1] == 3 and disp.dim() >= 4
            and disp.is_contiguous() and not _wants_grad(disp) and disp.numel() > 0
            and int(math.prod(disp.shape[-4:])) < 2 ** 31):
        X, Y, Z = disp.shape[-4:-1]
        Bn = int(math.prod(disp.shape[:-4]))
        out = torch.empty_like(disp)
        _lib.check(_lib.lib().bfm_compose_step(disp.data_ptr(), out.data_ptr(), Bn, X, Y, Z,
                                               (C.c_int * 3)(*_bounds(bound, 3)), _extrap(extrapolate), _stream()))
        return out
    dim = disp.shape[-1]
    moved = torch.movedim(disp, -1, -dim - 1)
    pulled = grid_pull(moved, add_identity_grid(disp), interpolation=1, bound=bound, extrapolate=extrapolate)
    return disp + torch.movedim(pulled, -dim - 1, -1)


def exp_velocity(svf, steps=7, bound='dct2', extrapolate=True):
    """Scaling and squaring: the displacement of exp(svf), `disp = svf / 2**steps`, then `steps` compose_step calls."""
    if (svf.is_cuda and svf.dtype == torch.float32 and svf.shape[-1] == 3 and svf.dim() >= 4 and svf.is_contiguous()
            and not _wants_grad(svf) and svf.numel() > 0 and 0 <= int(steps) < 31
            and int(math.prod(svf.shape[-4:-1])) * 4 < 2 ** 31):
        # one library call: the field stays in {x, y, z, 0} records between the steps (bfm_exp_velocity)
        X, Y, Z = svf.shape[-4:-1]
        Bn = int(math.prod(svf.shape[:-4]))
        out = torch.empty_like(svf)
        scratch = torch.empty(2 * Bn * X * Y * Z * 4, dtype=torch.float32, device=svf.device)
        _lib.check(_lib.lib().bfm_exp_velocity(svf.data_ptr(), out.data_ptr(), Bn, X, Y, Z, int(steps),
                                               (C.c_int * 3)(*_bounds(bound, 3)), _extrap(extrapolate),
                                               scratch.data_ptr(), _stream()))
        return out
    disp = svf / 2 ** steps
    for _ in range(steps):
        disp = compose_step(disp, bound=bound, extrapolate=extrapolate)
    return disp


def affine_grid(mat, shape):
    """Dense transformation grid from an affine matrix (api.py:524-560)."""
    mat = torch.as_tensor(mat)
    shape = list(shape)
    nb_dim = mat.shape[-1] - 1
    if nb_dim != len(shape):
        raise ValueError('Dimension of the affine matrix ({}) and shape ({}) are not the same.'
                         .format(nb_dim, len(shape)))
    if mat.shape[-2] not in (nb_dim, nb_dim + 1):
        raise ValueError('First argument should be matrces of shape (..., {0}, {1}) or (..., {1], {1}) but got {2}.'
                         .format(nb_dim, nb_dim + 1, mat.shape))
    batch_shape = mat.shape[:-2]
    grid = identity_grid(shape, mat.dtype, mat.device)
    if batch_shape:
        for _ in range(len(batch_shape)):
            grid = grid.unsqueeze(0)
        for _ in range(nb_dim):
            mat = mat.unsqueeze(-1)
    lin = mat[..., :nb_dim, :nb_dim]
    off = mat[..., :nb_dim, -1]
    grid = torch.matmul(lin, grid.unsqueeze(-1)).squeeze(-1) + off
    return grid


pull = grid_pull
push = grid_push
count = grid_count


def _make_list(x, n=None):
    x = list(x) if isinstance(x, (list, tuple)) else [x]
    if n is not None:
        x = (x + [x[-1]] * n)[:n]
    return x


def resize(image, factor=None, shape=None, anchor='c', interpolation=1, prefilter=True, **kwargs):
    """Resize an image by a factor or to a specific shape (utils/interpol/resize.py:13-119)."""
    factor = _make_list(factor) if factor else []
    shape = _make_list(shape) if shape else []
    anchor = _make_list(anchor)
    nb_dim = max(len(factor), len(shape), len(anchor)) or (image.dim() - 2)
    anchor = [a[0].lower() for a in _make_list(anchor, nb_dim)]
    bck = dict(dtype=image.dtype, device=image.device)
    inshape = image.shape[-nb_dim:]
    if factor:
        factor = _make_list(factor, nb_dim)
    elif not shape:
        raise ValueError('One of `factor` or `shape` must be provided')
    if shape:
        shape = _make_list(shape, nb_dim)
    else:
        shape = [int(i * f) for i, f in zip(inshape, factor)]
    if not factor:
        factor = [o / i for o, i in zip(shape, inshape)]
    lin = []
    for anch, f, inshp, outshp in zip(anchor, factor, inshape, shape):
        if anch == 'c':
            lin.append(torch.linspace(0, inshp - 1, outshp, **bck))
        elif anch == 'e':
            scale = inshp / outshp
            shift = 0.5 * (scale - 1)
            lin.append(torch.arange(0., outshp, **bck) * scale + shift)
        elif anch == 'f':
            lin.append(torch.arange(0., outshp, **bck) / f)
        elif anch == 'l':
            shift = (inshp - 1) - (outshp - 1) / f
            lin.append(torch.arange(0., outshp, **bck) / f + shift)
        else:
            raise ValueError('Unknown anchor {}'.format(anch))
    kwargs.setdefault('bound', 'nearest')
    kwargs.setdefault('extrapolate', True)
    kwargs.setdefault('interpolation', interpolation)
    kwargs.setdefault('prefilter', prefilter)
    grid = torch.stack(torch.meshgrid(*lin, indexing='ij'), dim=-1)
    return grid_pull(image, grid, **kwargs)


def restrict(image, factor=None, shape=None, anchor='c', interpolation=1, reduce_sum=False, **kwargs):
    """Restrict an image by a factor or to a specific shape: adjoint of resize (restrict.py:9-120)."""
    factor = _make_list(factor) if factor else []
    shape = _make_list(shape) if shape else []
    anchor = _make_list(anchor)
    nb_dim = max(len(factor), len(shape), len(anchor)) or (image.dim() - 2)
    anchor = [a[0].lower() for a in _make_list(anchor, nb_dim)]
    bck = dict(dtype=image.dtype, device=image.device)
    inshape = image.shape[-nb_dim:]
    if factor:
        factor = _make_list(factor, nb_dim)
    elif not shape:
        raise ValueError('One of `factor` or `shape` must be provided')
    if shape:
        shape = _make_list(shape, nb_dim)
    else:
        shape = [int(i / f) for i, f in zip(inshape, factor)]
    if not factor:
        factor = [i / o for o, i in zip(shape, inshape)]
    lin = []
    fullscale = 1
    for anch, f, inshp, outshp in zip(anchor, factor, inshape, shape):
        if anch == 'c':
            lin.append(torch.linspace(0, outshp - 1, inshp, **bck))
            fullscale *= (inshp - 1) / (outshp - 1)
        elif anch == 'e':
            scale = outshp / inshp
            shift = 0.5 * (scale - 1)
            fullscale *= scale
            lin.append(torch.arange(0., inshp, **bck) * scale + shift)
        elif anch == 'f':
            fullscale *= 1 / f
            lin.append(torch.arange(0., inshp, **bck) / f)
        elif anch == 'l':
            shift = (outshp - 1) - (inshp - 1) / f
            fullscale *= 1 / f
            lin.append(torch.arange(0., inshp, **bck) / f + shift)
        else:
            raise ValueError('Unknown anchor {}'.format(anch))
    kwargs.setdefault('bound', 'nearest')
    kwargs.setdefault('extrapolate', True)
    kwargs.setdefault('interpolation', interpolation)
    kwargs.setdefault('prefilter', False)
    grid = torch.stack(torch.meshgrid(*lin, indexing='ij'), dim=-1)
    resized = grid_push(image, grid, shape, **kwargs)
    if not reduce_sum:
        resized /= fullscale
    return resized
