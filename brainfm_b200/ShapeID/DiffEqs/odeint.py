"""ODE integration (mirror of ShapeID/DiffEqs/odeint.py:20-75, dopri5.py:58-175, rk_common.py:22-80,
interp.py:5-65, misc.py:84-170, fixed_grid.py, solvers.py:44-216).

The Runge-Kutta stage combinations, the error ratio and (for AdvDiffPDE) the right-hand side are libbfm
kernels; the step-size controller stays on the host with ONE 8-byte device->host read per step, and reproduces
the reference's non-textbook clamp / forced-accept logic (dopri5.py:152-169).
`solver.trace` records (t0, dt, accepted, ratio) per step and `solver.n_rhs` the number of RHS evaluations -- the
integer parity checks on control flow."""
import ctypes as C

import numpy as np
import torch

from ... import _lib
from .._common import need_cuda, stream

# Dormand-Prince (dopri5.py:11-36)
_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1., 1.]
_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
_C_ERROR = [35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
            -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1. / 60.]
_C_MID = [6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
          187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2]


def _f32(x):
    return float(torch.tensor(x, dtype=torch.float32))


def _combine(y0, ks, coefs, out_dtype=None):
    """y0 + sum_j float32(coef_j) * k_j  (float32 partial sums; y0 None -> the bare float32 sum)."""
    n = ks[0].numel()
    kp = (C.c_void_p * len(ks))(*[k.data_ptr() for k in ks])
    cf = (C.c_float * len(ks))(*[float(c) for c in coefs])
    if y0 is None:
        out = torch.empty(ks[0].shape, dtype=torch.float32, device=ks[0].device)
        _lib.check(_lib.lib().bfm_rk_combine(None, 0, kp, cf, len(ks), n, out.data_ptr(), 0, stream()))
        return out
    dbl = y0.dtype == torch.float64
    out = torch.empty_like(y0)
    _lib.check(_lib.lib().bfm_rk_combine(y0.data_ptr(), 1 if dbl else 0, kp, cf, len(ks), n, out.data_ptr(),
                                         1 if dbl else 0, stream()))
    return out


def _scaled(dt, coefs, dtype):
    """float32 value of (dt * c) as the reference forms it: dt is a 0-dim tensor of the state dtype, c a python
    float (cast to that dtype), the product rounded in that dtype and then to float32 -- evaluated with numpy
    scalars (same IEEE operations, no tensor construction per coefficient)."""
    T = np.float64 if dtype == torch.float64 else np.float32
    d = T(float(dt))
    return [float(np.float32(d * T(c))) for c in coefs]


def _rms(x):
    return float(x.norm() / (x.numel() ** 0.5))


class _Base:
    def __init__(self, func, y0, rtol, atol, dt, options=None):
        need_cuda(y0, "y0")
        if not torch.is_floating_point(y0):
            raise TypeError('`y0` must be a floating point Tensor but is a {}'.format(y0.type()))
        self.func, self.y0, self.rtol, self.atol, self.dt = func, y0.contiguous(), rtol, atol, dt
        self.n_rhs = 0
        self.trace = []

    def f(self, t, y):
        self.n_rhs += 1
        out = self.func(torch.as_tensor(t, dtype=y.dtype), y)
        if out.dtype != torch.float32:
            out = out.float()          # the kernels combine float32 stages (RHS of AdvDiffPDE is float32)
        return out.contiguous()


class Dopri5Solver(_Base):
    """Adaptive Dormand-Prince 5(4) with the reference's step clamp (dopri5.py:58-175)."""

    def _initial_step(self, t0, y0, f0):
        rtol, atol = self.rtol, self.atol
        scale = atol + torch.abs(y0) * rtol
        d0, d1 = _rms(y0 / scale), _rms(f0 / scale)
        h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * (d0 / d1)
        h0 = float(torch.tensor(h0, dtype=y0.dtype))
        y1 = y0 + h0 * f0
        f1 = self.f(t0 + h0, y1)
        d2 = _rms((f1 - f0) / scale) / h0
        if d1 <= 1e-15 and d2 <= 1e-15:
            h1 = max(1e-6, h0 * 1e-3)
        else:
            h1 = (0.01 / max(d1, d2)) ** (1. / 5.)     # `max(d1 + d2)` is a tuple concatenation (misc.py:141)
        return min(100 * h0, h1)

    def _step(self, y0, f0, t0, dt):
        dt_state = float(torch.tensor(dt, dtype=y0.dtype))
        ks = [f0]
        yi = y0
        for alpha, beta in zip(_ALPHA, _BETA):
            yi = _combine(y0, ks, _scaled(dt_state, beta, y0.dtype))
            ks.append(self.f(t0 + alpha * dt_state, yi))
        y1 = yi
        # error estimate (float32 sum of the stages) and its scaled square sum in one pass; err is never stored
        n = y0.numel()
        kp = (C.c_void_p * len(ks))(*[k.data_ptr() for k in ks])
        cf = (C.c_float * len(ks))(*_scaled(dt_state, _C_ERROR, y0.dtype))
        acc = torch.empty(1, dtype=torch.float64, device=y0.device)
        aligned = n % 4 == 0 and all(k.data_ptr() % 16 == 0 for k in ks) and y0.data_ptr() % 16 == 0 \
            and y1.data_ptr() % 16 == 0
        scratch = None if aligned else torch.empty(n, dtype=torch.float32, device=y0.device)
        _lib.check(_lib.lib().bfm_rk_error_fused(kp, cf, len(ks), y0.data_ptr(), y1.data_ptr(),
                                                 1 if y0.dtype == torch.float64 else 0, n, float(self.rtol),
                                                 float(self.atol), None if scratch is None else scratch.data_ptr(),
                                                 acc.data_ptr(), stream()))
        ratio = float(acc.item()) / y0.numel()      # the one device->host read of the step
        return y1, ks, ratio, dt_state

    @staticmethod
    def _optimal_step(last, ratio, safety=0.9, ifactor=10.0, dfactor=0.2, order=5):
        if ratio == 0:
            return last * ifactor
        if ratio < 1:
            dfactor = 1.0
        factor = max(1 / ifactor, min((ratio ** 0.5) ** (1 / order) / safety, 1 / dfactor))
        return last / factor

    def _interp(self, state, t):
        """Quartic dense output through y0, y1, y_mid, f0, f1 (interp.py:5-65), fitted lazily."""
        y0, y1, ks, dt, t0, t1 = state
        dtt = float(torch.tensor(dt, dtype=y0.dtype))
        y_mid = _combine(y0, ks, _scaled(dtt, _C_MID, y0.dtype))
        f0, f1 = ks[0], ks[-1]
        a = (-2 * dtt) * f0 + (2 * dtt) * f1 + -8 * y0 + -8 * y1 + 16 * y_mid
        b = (5 * dtt) * f0 + (-3 * dtt) * f1 + 18 * y0 + 14 * y1 + -32 * y_mid
        c = (-4 * dtt) * f0 + dtt * f1 + -11 * y0 + -5 * y1 + 16 * y_mid
        d = dtt * f0
        e = y0
        T = y0.dtype
        tt0, tt1, tt = (float(torch.tensor(v, dtype=T)) for v in (t0, t1, t))
        assert tt0 <= tt <= tt1, 'invalid interpolation, fails `t0 <= t <= t1`: {}, {}, {}'.format(tt0, tt, tt1)
        x = torch.tensor((tt - tt0) / (tt1 - tt0), dtype=T)
        x2 = x * x
        x3 = x2 * x
        x4 = x3 * x
        return a * x4 + b * x3 + c * x2 + d * x + e * torch.tensor(1, dtype=T)

    def integrate(self, t):
        t = [float(v) for v in t.to(torch.float64).cpu()]
        assert all(b > a for a, b in zip(t[:-1], t[1:])), 't must be strictly increasing or decrasing'
        y0 = self.y0
        f0 = self.f(t[0], y0)
        dt = self._initial_step(t[0], y0, f0)
        t0 = t1 = t[0]
        last = None
        solution = [y0]
        tol_min_dt = 0.2 * self.dt if 0.1 * self.dt >= 0.01 else 0.01
        for next_t in t[1:]:
            n_steps = 0
            while next_t > t1:
                assert n_steps < 2 ** 31 - 1, 'max_num_steps exceeded'
                assert t1 + dt > t1, 'underflow in dt {}'.format(dt)
                y1, ks, ratio, dt_state = self._step(y0, f0, t1, dt)
                accept = ratio <= 1
                dt_next = self._optimal_step(dt, ratio)
                forced = dt_next < tol_min_dt or dt_next > 0.1
                if forced:
                    if dt_next < tol_min_dt:
                        dt_next = tol_min_dt
                    if dt_next > 0.1:
                        dt_next = 0.1
                self.trace.append((t1, dt, bool(accept or forced), ratio))
                if accept or forced:
                    last = (y0, y1, ks, dt, t1, t1 + dt)
                    t0, t1 = t1, t1 + dt
                    y0, f0 = y1, ks[-1]
                dt = dt_next
                n_steps += 1
            solution.append(self._interp(last, next_t) if last is not None else y0)
        return torch.stack(solution)


class _FixedGrid(_Base):
    """Fixed time grid = the requested output times (solvers.py:103-216)."""

    def step(self, t, dt, y):
        raise NotImplementedError

    def integrate(self, t):
        tt = t.to(self.y0.dtype)
        ts = [float(v) for v in tt.cpu()]
        assert all(b > a for a, b in zip(ts[:-1], ts[1:])), 't must be strictly increasing or decrasing'
        y = self.y0
        solution = [y]
        for t0, t1 in zip(ts[:-1], ts[1:]):
            dt = float(torch.tensor(t1, dtype=y.dtype) - torch.tensor(t0, dtype=y.dtype))
            y = self.step(t0, dt, y)
            solution.append(y)
        return torch.stack(solution)


class Euler(_FixedGrid):
    order = 1

    def step(self, t, dt, y):                       # dy = dt * f (fixed_grid.py:5-12)
        return _combine(y, [self.f(t, y)], [_f32(dt)])


class Midpoint(_FixedGrid):
    order = 2

    def step(self, t, dt, y):                       # fixed_grid.py:15-23
        k1 = self.f(t, y)
        y_mid = y + k1 * dt / 2
        return _combine(y, [self.f(t + dt / 2, y_mid)], [_f32(dt)])


class RK4(_FixedGrid):
    order = 4

    def step(self, t, dt, y):                       # rk4_alt_step_func (rk_common.py:72-79)
        k1 = self.f(t, y)
        k2 = self.f(t + dt / 3, y + dt * k1 / 3)
        k3 = self.f(t + dt * 2 / 3, y + dt * (k1 / -3 + k2))
        k4 = self.f(t + dt, y + dt * (k1 - k2 + k3))
        return y + (k1 + 3 * k2 + 3 * k3 + k4) * (dt / 8)


def _unbuilt(name):
    class _Missing:
        def __init__(self, *a, **k):
            raise NotImplementedError("ODE method %r is not built yet (SURVEY.md 8f-4)" % name)
    return _Missing


SOLVERS = {
    'explicit_adams': _unbuilt('explicit_adams'),
    'fixed_adams': _unbuilt('fixed_adams'),
    'adams': _unbuilt('adams'),
    'tsit5': _unbuilt('tsit5'),
    'dopri5': Dopri5Solver,
    'euler': Euler,
    'midpoint': Midpoint,
    'rk4': RK4,
}


def odeint(func, y0, t, dt, step_size=None, rtol=1e-7, atol=1e-9, method=None, options=None, return_solver=False):
    """Integrate dy/dt = func(t, y), y(t[0]) = y0 (ShapeID/DiffEqs/odeint.py:20-75).  y0: one tensor (or a
    1-tuple); returns the solution stacked along a new first axis."""
    tensor_input = torch.is_tensor(y0)
    if not tensor_input:
        assert isinstance(y0, tuple), 'y0 must be either a torch.Tensor or a tuple'
        if len(y0) != 1:
            raise NotImplementedError("tuple states with more than one tensor are not supported")
        base = func
        func = lambda tt, y: base(tt, (y,))[0]      # noqa: E731
        y0 = y0[0]
    if not torch.is_floating_point(t):
        raise TypeError('`t` must be a floating point Tensor but is a {}'.format(t.type()))
    if options and method is None:
        raise ValueError('cannot supply `options` without specifying `method`')
    if method is None:
        method = 'dopri5'
    if method not in SOLVERS:
        raise KeyError(method)
    if len(t) > 1 and bool((t[1:] < t[:-1]).all()):
        t = -t
        fwd = func
        func = lambda tt, y: -fwd(-tt, y)           # noqa: E731
    solver = SOLVERS[method](func, y0, rtol=rtol, atol=atol, dt=dt, options=options)
    sol = solver.integrate(t)
    out = sol if tensor_input else (sol,)
    return (out, solver) if return_solver else out
