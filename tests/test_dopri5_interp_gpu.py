"""bfm_dopri5_interp (the fused dense output of dopri5) against the tensor expressions it replaces
(brainfm_b200/ShapeID/DiffEqs/odeint.py::_interp, reference ShapeID/DiffEqs/interp.py:5-65): bit for bit."""
import ctypes as C

import pytest
import torch

from brainfm_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("x", [0.0, 0.37, 1.0])
def test_fused_dense_output_equals_the_tensor_expressions(x):
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(5)
    n = 40007
    y0, y1, ym = (torch.rand(n, generator=g, dtype=torch.float64).to(dev) for _ in range(3))
    f0, f1 = (torch.randn(n, generator=g, dtype=torch.float32).to(dev) for _ in range(2))
    dtt = 0.0731
    T = torch.float64
    a = (-2 * dtt) * f0 + (2 * dtt) * f1 + -8 * y0 + -8 * y1 + 16 * ym
    b = (5 * dtt) * f0 + (-3 * dtt) * f1 + 18 * y0 + 14 * y1 + -32 * ym
    c = (-4 * dtt) * f0 + dtt * f1 + -11 * y0 + -5 * y1 + 16 * ym
    d = dtt * f0
    e = y0
    xt = torch.tensor(x, dtype=T)
    x2 = xt * xt
    x3 = x2 * xt
    x4 = x3 * xt
    ref = a * x4 + b * x3 + c * x2 + d * xt + e * torch.tensor(1, dtype=T)
    out = torch.empty_like(y0)
    _lib.check(_lib.lib().bfm_dopri5_interp(y0.data_ptr(), y1.data_ptr(), ym.data_ptr(), f0.data_ptr(), f1.data_ptr(),
                                            dtt, x, n, out.data_ptr(),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    assert ref.dtype == torch.float64
    assert torch.equal(out, ref)
