"""odeint_adjoint (mirror of ShapeID/DiffEqs/adjoint.py:105-132).  The generator only runs the forward pass,
under no_grad; the adjoint backward is not built."""
import torch
import torch.nn as nn

from .odeint import odeint


def odeint_adjoint(func, y0, t, dt, rtol=1e-6, atol=1e-12, method=None, options=None, return_solver=False):
    if not isinstance(func, nn.Module):
        raise ValueError('func is required to be an instance of nn.Module.')
    if torch.is_tensor(y0) and y0.requires_grad:
        raise NotImplementedError("the adjoint backward pass is not built; call under torch.no_grad()")
    with torch.no_grad():
        return odeint(func, y0, t, dt, rtol=rtol, atol=atol, method=method, options=options,
                      return_solver=return_solver)
