"""Gradient checks of the CUDA interpol backend: the reference's own suite
(utils/interpol/tests/test_gradcheck_pushpull.py: float64, 3 samples per axis, 2 batch elements, extrapolate=True,
every bound for orders 0-2, dct2 for orders 3-7, 1-D / 2-D / 3-D) run against brainfm_b200.interpol for
grid_pull, grid_push, grid_count and the spline prefilter."""
import inspect

import pytest
import torch
from torch.autograd import gradcheck

pytestmark = pytest.mark.gpu

dtype = torch.double
shape1 = 3
extrapolate = True
kwargs = dict(rtol=1., raise_exception=True)
if 'check_undefined_grad' in inspect.signature(gradcheck).parameters:
    kwargs['check_undefined_grad'] = False
if 'nondet_tol' in inspect.signature(gradcheck).parameters:
    kwargs['nondet_tol'] = 1e-3

order_bounds = [(o, b) for o in range(3) for b in range(7)] + [(o, 3) for o in range(3, 8)]
NAMES = ['zero', 'replicate', 'dct1', 'dct2', 'dst1', 'dst2', 'dft']


def make_data(shape, seed):
    from brainfm_b200.interpol import add_identity_grid_
    g = torch.Generator(device='cpu').manual_seed(seed)
    grid = torch.randn([2, *shape, len(shape)], dtype=dtype, generator=g).cuda()
    grid = add_identity_grid_(grid)
    vol = torch.randn((2, 1,) + shape, dtype=dtype, generator=g).cuda()
    return vol, grid


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("interpolation,bound", order_bounds)
def test_gradcheck_pull(dim, bound, interpolation):
    from brainfm_b200.interpol import grid_pull
    vol, grid = make_data((shape1,) * dim, 100 * dim + 10 * interpolation + bound)
    vol.requires_grad = True
    grid.requires_grad = True
    assert gradcheck(grid_pull, (vol, grid, interpolation, NAMES[bound], extrapolate), **kwargs)


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("interpolation,bound", order_bounds)
def test_gradcheck_grad(dim, bound, interpolation):
    """utils/interpol/tests/test_gradcheck_pushpull.py:62-74: the backward of grid_grad (push with derivative weights
    + spline Hessians, bfm_interpol_grad_backward)."""
    from brainfm_b200.interpol import grid_grad
    vol, grid = make_data((shape1,) * dim, 300 * dim + 10 * interpolation + bound)
    vol.requires_grad = True
    grid.requires_grad = True
    assert gradcheck(grid_grad, (vol, grid, interpolation, NAMES[bound], extrapolate), **kwargs)


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("interpolation,bound", order_bounds)
def test_gradcheck_push(dim, bound, interpolation):
    from brainfm_b200.interpol import grid_push
    shape = (shape1,) * dim
    vol, grid = make_data(shape, 200 * dim + 10 * interpolation + bound)
    vol.requires_grad = True
    grid.requires_grad = True
    assert gradcheck(grid_push, (vol, grid, shape, interpolation, NAMES[bound], extrapolate), **kwargs)


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("interpolation,bound", order_bounds)
def test_gradcheck_count(dim, bound, interpolation):
    from brainfm_b200.interpol import grid_count
    shape = (shape1,) * dim
    _, grid = make_data(shape, 300 * dim + 10 * interpolation + bound)
    grid.requires_grad = True
    assert gradcheck(grid_count, (grid, shape, interpolation, NAMES[bound], extrapolate), **kwargs)


def test_adjoints_and_prefilter_backward():
    """<pull(x), y> == <x, push(y)> in float64; pull's input gradient IS push; the prefilter's backward is the
    prefilter; tighter gradcheck (default tolerances) of a smooth cubic pull away from the kinks."""
    from brainfm_b200 import interpol
    torch.manual_seed(0)
    x = torch.randn(2, 3, 9, 8, 7, dtype=dtype, device='cuda')
    grid = interpol.add_identity_grid_(0.7 * torch.randn(2, 6, 5, 4, 3, dtype=dtype, device='cuda'))
    y = torch.randn(2, 3, 6, 5, 4, dtype=dtype, device='cuda')
    for order, bound in ((1, 'dct2'), (3, 'dct2'), (2, 'dft'), (3, 'zero')):
        a = (interpol.grid_pull(x, grid, order, bound, True) * y).sum()
        b = (x * interpol.grid_push(y, grid, (9, 8, 7), order, bound, True)).sum()
        assert abs(float(a - b)) < 1e-9 * max(1.0, abs(float(a)))
        xr = x.clone().requires_grad_(True)
        (interpol.grid_pull(xr, grid, order, bound, True) * y).sum().backward()
        assert torch.allclose(xr.grad, interpol.grid_push(y, grid, (9, 8, 7), order, bound, True), atol=1e-12)
    xs = torch.randn(1, 2, 6, 6, 6, dtype=dtype, device='cuda', requires_grad=True)
    gs = (interpol.identity_grid([4, 4, 4], dtype=dtype, device='cuda')[None] + 1.3).requires_grad_(True)
    assert gradcheck(lambda a, g: interpol.grid_pull(a, g, 3, 'dct2', True, prefilter=True), (xs, gs))
    c = torch.randn(2, 10, 11, dtype=dtype, device='cuda', requires_grad=True)
    assert gradcheck(lambda a: interpol.spline_coeff_nd(a, 3, 'dct2', dim=2), (c,))
    # grid_grad is differentiable too (round 2: pushgrad + spline Hessians), prefilter included
    assert gradcheck(lambda a, g: interpol.grid_grad(a, g, 3, 'dct2', True, prefilter=True), (xs, gs))
