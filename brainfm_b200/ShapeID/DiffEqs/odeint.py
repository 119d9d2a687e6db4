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
        if (y0.dtype == torch.float64 and f0.dtype == torch.float32 and y0.is_contiguous() and y1.is_contiguous()
                and f0.is_contiguous() and f1.is_contiguous()):
            # one pass instead of fifteen element-wise tensor expressions (bfm_dopri5_interp: same operations, same
            # dtype promotions)
            tt0, tt1, tt = float(t0), float(t1), float(t)
            assert tt0 <= tt <= tt1, 'invalid interpolation, fails `t0 <= t <= t1`: {}, {}, {}'.format(tt0, tt, tt1)
            out = torch.empty_like(y0)
            _lib.check(_lib.lib().bfm_dopri5_interp(y0.data_ptr(), y1.data_ptr(), y_mid.data_ptr(), f0.data_ptr(),
                                                    f1.data_ptr(), dtt, (tt - tt0) / (tt1 - tt0), y0.numel(),
                                                    out.data_ptr(), stream()))
            return out
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


# ---------------------------------------------------------------------------------------------------------------
# Tsitouras 5(4) (tsit5.py).  Reproduced AS THE REFERENCE BEHAVES, not as the textbook method: its error estimate
# (c_error, tsit5.py:19-27) rejects steps until dt ~ 1e-7 on ordinary problems (the reference needed 300 s for a
# 2-element linear ODE on [0, 1]), and its dense output starts from k[0] = f0 instead of y0 (tsit5.py:45-50).
# The (t0, dt, accepted) trace of the first steps is pinned on the reference's (tests/golden/solvers.npz).
_TSIT_ALPHA = [0.161, 0.327, 0.9, 0.9800255409045097, 1., 1.]
_TSIT_BETA = [
    [0.161],
    [-0.008480655492357, 0.3354806554923570],
    [2.897153057105494, -6.359448489975075, 4.362295432869581],
    [5.32586482843925895, -11.74888356406283, 7.495539342889836, -0.09249506636175525],
    [5.86145544294642038, -12.92096931784711, 8.159367898576159, -0.071584973281401006, -0.02826905039406838],
    [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774],
]
_TSIT_C_ERROR = [
    0.09646076681806523 - 0.001780011052226, 0.01 - 0.000816434459657, 0.4798896504144996 - -0.007880878010262,
    1.379008574103742 - 0.144711007173263, -3.290069515436081 - -0.582357165452555,
    2.324710524099774 - 0.458082105929187, -1 / 66,
]


def _tsit5_interp_coeff(t0, dt, eval_t):
    t = float((eval_t - t0) / dt)
    b1 = -1.0530884977290216 * t * (t - 1.3299890189751412) * (t ** 2 - 1.4364028541716351 * t + 0.7139816917074209)
    b2 = 0.1017 * t ** 2 * (t ** 2 - 2.1966568338249754 * t + 1.2949852507374631)
    b3 = 2.490627285651252793 * t ** 2 * (t ** 2 - 2.38535645472061657 * t + 1.57803468208092486)
    b4 = -16.54810288924490272 * (t - 1.21712927295533244) * (t - 0.61620406037800089) * t ** 2
    b5 = 47.37952196281928122 * (t - 1.203071208372362603) * (t - 0.658047292653547382) * t ** 2
    b6 = -34.87065786149660974 * (t - 1.2) * (t - 0.666666666666666667) * t ** 2
    b7 = 2.5 * (t - 1) * (t - 0.6) * t ** 2
    return [b1, b2, b3, b4, b5, b6, b7]


class Tsit5Solver(Dopri5Solver):
    """tsit5.py:60-139: the Dormand-Prince machinery with the Tsitouras tableau, the plain step-size controller (no
    clamp, no forced accept) and the reference's dense output.  `max_num_steps` bounds the steps per output time
    (AssertionError like the reference's)."""

    def __init__(self, func, y0, rtol, atol, dt=None, options=None, max_num_steps=2 ** 31 - 1):
        super(Tsit5Solver, self).__init__(func, y0, rtol, atol, dt, options)
        self.max_num_steps = (options or {}).get('max_num_steps', max_num_steps) if isinstance(options, dict) \
            else max_num_steps

    def _rk_step(self, y0, f0, t0, dt):
        dt_state = float(torch.tensor(dt, dtype=y0.dtype))
        ks = [f0]
        yi = y0
        for alpha, beta in zip(_TSIT_ALPHA, _TSIT_BETA):
            yi = _combine(y0, ks, _scaled(dt_state, beta, y0.dtype))
            ks.append(self.f(t0 + alpha * dt_state, yi))
        y1 = yi                                   # c_sol == the last beta row (+ 0 * k7)
        err = _combine(None, ks, _scaled(dt_state, _TSIT_C_ERROR, y0.dtype))
        n = y0.numel()
        acc = torch.empty(1, dtype=torch.float64, device=y0.device)
        _lib.check(_lib.lib().bfm_rk_error_sum(err.data_ptr(), y0.data_ptr(), y1.data_ptr(),
                                               1 if y0.dtype == torch.float64 else 0, n, float(self.rtol),
                                               float(self.atol), acc.data_ptr(), stream()))
        return y1, ks, float(acc.item()) / n

    def integrate(self, t):
        t = [float(v) for v in t.to(torch.float64).cpu()]
        assert all(b > a for a, b in zip(t[:-1], t[1:])), 't must be strictly increasing or decrasing'
        y0 = self.y0
        f0 = self.f(t[0], y0)                      # _select_initial_step evaluates f(t0, y0) itself (no f0 passed,
        dt = self._initial_step(t[0], y0, f0)      # tsit5.py:76) and before_integrate evaluates it again (:81):
        f0 = self.f(t[0], y0)                      # three RHS evaluations before the first step, like the reference
        t0 = t1 = t[0]
        ks = [y0] * 7                             # interp_coeff before any accepted step (tsit5.py:80-83)
        solution = [y0]
        for next_t in t[1:]:
            n_steps = 0
            while next_t > t1:
                assert n_steps < self.max_num_steps, 'max_num_steps exceeded ({}>={})'.format(n_steps, self.max_num_steps)
                assert t1 + dt > t1, 'underflow in dt {}'.format(dt)
                y1, k_new, ratio = self._rk_step(y0, f0, t1, dt)
                assert ratio == ratio, 'non-finite values in state `y`'
                accept = ratio <= 1
                self.trace.append((t1, dt, bool(accept), ratio))
                t0 = t1                            # rk_state.t0 is the step's start even when it is rejected (:138)
                if accept:
                    t1 = t1 + dt
                    y0, f0, ks = y1, k_new[-1], k_new
                dt = self._optimal_step(dt, ratio)
                n_steps += 1
            coef = _tsit5_interp_coeff(t0, t1 - t0, next_t) if t1 != t0 else [0.0] * 7
            T = np.float64 if y0.dtype == torch.float64 else np.float32
            dtt = T(t1 - t0)
            cf = [float(T(dtt * T(c))) for c in coef]
            if y0.dtype == torch.float32:
                solution.append(_combine(ks[0].to(y0.dtype), [k.float() for k in ks], cf))
            else:
                solution.append(ks[0].to(y0.dtype) + sum(c * k.to(y0.dtype) for c, k in zip(cf, ks)))
        return torch.stack(solution)


# ---------------------------------------------------------------------------------------------------------------
# Variable-coefficient Adams-Bashforth-Moulton (adams.py; Hairer, Norsett, Wanner III.5)
_GAMMA_STAR = [1, -1 / 2, -1 / 12, -1 / 24, -19 / 720, -3 / 160, -863 / 60480, -275 / 24192, -33953 / 3628800,
               -0.00789255, -0.00678585, -0.00592406, -0.00523669, -0.0046775, -0.00421495, -0.0038269]


def _lincomb(y0, tensors, coefs):
    """y0 + sum_j coef_j * tensor_j in the state dtype (y0 None: the bare sum): the RK combination kernel for float32
    states, torch arithmetic for float64 ones (the kernel sums float32 partials)."""
    if not tensors:
        return y0
    ref = y0 if y0 is not None else tensors[0]
    if ref.dtype == torch.float32 and all(t.dtype == torch.float32 for t in tensors):
        return _combine(y0, [t.contiguous() for t in tensors], [_f32(c) for c in coefs])
    acc = None
    for c, t in zip(coefs, tensors):
        term = t.to(ref.dtype) * c
        acc = term if acc is None else acc + term
    return acc if y0 is None else y0 + acc


def _error_ratio(err, y0, y1, rtol, atol):
    """mean((err / (atol + rtol * max(|y0|, |y1|)))^2)  (misc.py:146-157) -- one reduction kernel, one 8-byte read."""
    n = y0.numel()
    acc = torch.empty(1, dtype=torch.float64, device=y0.device)
    e32 = err.float().contiguous()
    _lib.check(_lib.lib().bfm_rk_error_sum(e32.data_ptr(), y0.data_ptr(), y1.data_ptr(),
                                           1 if y0.dtype == torch.float64 else 0, n, float(rtol), float(atol),
                                           acc.data_ptr(), stream()))
    return float(acc.item()) / n


class VariableCoefficientAdamsBashforth(_Base):
    """adams.py:59-170.  Orders 1..12, predictor-corrector with the implicit (modified divided difference) update,
    order and step-size selection as in the reference.  `trace` rows: (t_n, dt, accepted, error ratio, order)."""
    _MAX_ORDER = 12

    def __init__(self, func, y0, rtol, atol, dt=None, options=None, implicit=True, max_order=12, safety=0.9,
                 ifactor=10.0, dfactor=0.2):
        super(VariableCoefficientAdamsBashforth, self).__init__(func, y0, rtol, atol, dt, options)
        self.implicit = implicit
        self.max_order = int(max(1, min(max_order, self._MAX_ORDER)))
        self.safety, self.ifactor, self.dfactor = safety, ifactor, dfactor

    def _initial_step2(self, t0, y0, f0):
        """_select_initial_step(order=2, f0 given)  (misc.py:84-143)."""
        rtol, atol = self.rtol, self.atol
        scale = atol + torch.abs(y0) * rtol
        d0, d1 = _rms(y0 / scale), _rms(f0.to(y0.dtype) / scale)
        h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * (d0 / d1)
        h0 = float(torch.tensor(h0, dtype=y0.dtype))
        y1 = y0 + h0 * f0.to(y0.dtype)
        f1 = self.f(t0 + h0, y1)
        d2 = _rms((f1 - f0).to(y0.dtype) / scale) / h0
        if d1 <= 1e-15 and d2 <= 1e-15:
            h1 = max(1e-6, h0 * 1e-3)
        else:
            h1 = (0.01 / max(d1, d2)) ** (1. / 3.)
        return min(100 * h0, h1)

    @staticmethod
    def _g_and_beta(prev_t, next_t, k):
        """g[0..k] and the beta_j that turn implicit phi_j into explicit ones (adams.py:26-48), float64 scalars."""
        curr_t = prev_t[0]
        dt = next_t - prev_t[0]
        g = [0.0] * (k + 1)
        betas = [1.0]
        beta = 1.0
        g[0] = 1.0
        c = [1.0 / q for q in range(1, k + 2)]
        for j in range(1, k):
            beta = (next_t - prev_t[j - 1]) / (curr_t - prev_t[j]) * beta
            betas.append(beta)
            if j == 1:
                c = [a - b for a, b in zip(c[:-1], c[1:])]
            else:
                c = [a - b * dt / (next_t - prev_t[j - 1]) for a, b in zip(c[:-1], c[1:])]
            g[j] = c[0]
        c = [a - b * dt / (next_t - prev_t[k - 1]) for a, b in zip(c[:-1], c[1:])]
        g[k] = c[0]
        return g, betas

    @staticmethod
    def _implicit_phi(explicit_phi, f_n, k):
        k = min(len(explicit_phi) + 1, k)
        out = [f_n]
        for j in range(1, k):
            out.append(out[j - 1] - explicit_phi[j - 1])
        return out

    def _opt_step(self, last, ratio, order):
        if ratio == 0:
            return last * self.ifactor
        dfactor = 1.0 if ratio < 1 else self.dfactor
        factor = max(1 / self.ifactor, min((ratio ** 0.5) ** (1 / order) / self.safety, 1 / dfactor))
        return last / factor

    def integrate(self, t):
        ts = [float(v) for v in t.to(torch.float64).cpu()]
        assert all(b > a for a, b in zip(ts[:-1], ts[1:])), 't must be strictly increasing or decrasing'
        import collections
        y_n = self.y0
        prev_t = collections.deque(maxlen=self.max_order + 1)
        phi = collections.deque(maxlen=self.max_order)
        f0 = self.f(ts[0], y_n)
        prev_t.appendleft(ts[0])
        phi.appendleft(f0)
        phi = list(phi)
        next_t = ts[0] + self._initial_step2(ts[0], y_n, f0)
        order = 1
        solution = [y_n]
        T = np.float64 if y_n.dtype == torch.float64 else np.float32
        for final_t in ts[1:]:
            while final_t > prev_t[0]:
                nt = min(next_t, final_t)
                dt = nt - prev_t[0]
                dt_c = float(T(dt))
                g, betas = self._g_and_beta(list(prev_t), nt, order)
                g = [float(T(v)) for v in g]
                ephi = [phi[0]] + [phi[j] * float(T(betas[j])) for j in range(1, order)]
                m = max(1, order - 1)
                p_next = _lincomb(y_n, ephi[:m], [float(T(dt_c * g[j])) for j in range(m)])
                next_f0 = self.f(nt, p_next)
                iphi_p = self._implicit_phi(ephi, next_f0, order + 1)
                y_next = _lincomb(p_next, [iphi_p[order - 1]], [float(T(T(dt_c * g[order - 1])))])
                local_error = iphi_p[order] * float(T(dt_c * (g[order] - g[order - 1])))
                error_k = _error_ratio(local_error, y_n, y_next, self.rtol, self.atol)
                assert error_k == error_k, 'non-finite values in state `y`'
                assert prev_t[0] + dt > prev_t[0], 'underflow in dt {}'.format(dt)
                accept = error_k <= 1
                self.trace.append((prev_t[0], dt, bool(accept), error_k, order))
                if not accept:
                    next_t = prev_t[0] + self._opt_step(dt, error_k, order)
                    continue
                next_f0 = self.f(nt, y_next)
                iphi = self._implicit_phi(ephi, next_f0, order + 2)
                next_order = order
                if len(prev_t) <= 4 or order < 3:
                    next_order = min(order + 1, 3, self.max_order)
                else:
                    e1 = _error_ratio(iphi_p[order - 1] * float(T(dt_c * (g[order - 1] - g[order - 2]))), y_n, y_next,
                                      self.rtol, self.atol)
                    e2 = _error_ratio(iphi_p[order - 2] * float(T(dt_c * (g[order - 2] - g[order - 3]))), y_n, y_next,
                                      self.rtol, self.atol)
                    if min(e1, e2) < error_k:      # `min(error_km1 + error_km2)`: the + concatenates two 1-tuples (adams.py:152)
                        next_order = order - 1
                    elif order < self.max_order:
                        e3 = _error_ratio(iphi_p[order] * float(T(dt_c * _GAMMA_STAR[order])), y_n, y_next, self.rtol,
                                          self.atol)
                        if e3 < error_k:
                            next_order = order + 1
                dt_next = dt if next_order > order else self._opt_step(dt, error_k, order + 1)
                prev_t.appendleft(nt)
                # the reference continues from the PREDICTOR p_next, not from the corrected y_next (adams.py:169)
                y_n, phi, order = p_next, iphi[:self.max_order], next_order
                next_t = nt + dt_next
            assert final_t == prev_t[0]
            solution.append(y_n)
        return torch.stack(solution)


# ---------------------------------------------------------------------------------------------------------------
# Fixed-grid Adams-Bashforth(-Moulton) (fixed_adams.py).  The reference's step_func calls `rk_common.rk4_alt_step_func`
# through a name its module never binds (`import ShapeID.DiffEqs.rk_common`, fixed_adams.py:5,165) and therefore
# raises NameError on the first step; this is the algorithm of that file with the name resolved (the golden vectors
# come from the reference with exactly that one attribute supplied, oracle/make_golden_solvers.py).
_BASHFORTH = {
    1: ([11], 11), 2: ([3, -1], 2), 3: ([23, -16, 5], 12), 4: ([55, -59, 37, -9], 24),
    5: ([1901, -2774, 2616, -1274, 251], 720), 6: ([4277, -7923, 9982, -7298, 2877, -475], 1440),
    7: ([198721, -447288, 705549, -688256, 407139, -134472, 19087], 60480),
    8: ([434241, -1152169, 2183877, -2664477, 2102243, -1041723, 295767, -36799], 120960),
    9: ([14097247, -43125206, 95476786, -139855262, 137968480, -91172642, 38833486, -9664106, 1070017], 3628800),
    10: ([30277247, -104995189, 265932680, -454661776, 538363838, -444772162, 252618224, -94307320, 20884811,
          -2082753], 7257600),
    11: ([2132509567, -8271795124, 23591063805, -46113029016, 63716378958, -63176201472, 44857168434, -22329634920,
          7417904451, -1479574348, 134211265], 479001600),
}
_MOULTON = {
    1: ([1], 11), 2: ([1, 1], 2), 3: ([5, 8, -1], 12), 4: ([9, 19, -5, 1], 24), 5: ([251, 646, -264, 106, -19], 720),
    6: ([475, 1427, -798, 482, -173, 27], 1440), 7: ([19087, 65112, -46461, 37504, -20211, 6312, -863], 60480),
    8: ([36799, 139849, -121797, 123133, -88547, 41499, -11351, 1375], 120960),
    9: ([1070017, 4467094, -4604594, 5595358, -5033120, 3146338, -1291214, 312874, -33953], 3628800),
    10: ([2082753, 9449717, -11271304, 16002320, -17283646, 13510082, -7394032, 2687864, -583435, 57281], 7257600),
    11: ([134211265, 656185652, -890175549, 1446205080, -1823311566, 1710774528, -1170597042, 567450984, -184776195,
          36284876, -3250433], 479001600),
    12: ([262747265, 1374799219, -2092490673, 3828828885, -5519460582, 6043521486, -4963166514, 3007739418,
          -1305971115, 384709327, -68928781, 5675265], 958003200),
}


class AdamsBashforthMoulton(_FixedGrid):
    """fixed_adams.py:136-205: RK4 (3/8 rule) until three derivatives are known, then Adams-Bashforth of order up to
    11 with an Adams-Moulton corrector iterated to convergence (at most max_iters times)."""
    order = 4

    def __init__(self, func, y0, rtol=1e-3, atol=1e-4, dt=None, options=None, implicit=True, max_iters=4, max_order=12):
        super(AdamsBashforthMoulton, self).__init__(func, y0, rtol, atol, dt, options)
        self.implicit, self.max_iters = implicit, max_iters
        self.max_order = int(min(max_order, 12))
        import collections
        self.prev_f = collections.deque(maxlen=self.max_order - 1)
        self.prev_t = None

    def _update_history(self, t, f):
        if self.prev_t is None or self.prev_t != t:
            self.prev_f.appendleft(f)
            self.prev_t = t

    def _converged(self, a, b):
        tol = self.atol + self.rtol * torch.max(torch.abs(a), torch.abs(b))
        return bool((torch.abs(a - b) < tol).all())

    def step(self, t, dt, y):
        self._update_history(t, self.f(t, y))
        order = min(len(self.prev_f), self.max_order - 1)
        if order < 3:
            k1 = self.prev_f[0]                                  # rk4_alt_step_func(func, t, dt, y, k1=prev_f[0])
            k2 = self.f(t + dt / 3, y + dt * k1 / 3)
            k3 = self.f(t + dt * 2 / 3, y + dt * (k1 / -3 + k2))
            k4 = self.f(t + dt, y + dt * (k1 - k2 + k3))
            return y + (k1 + 3 * k2 + 3 * k3 + k4) * (dt / 8)
        coefs, div = _BASHFORTH[order]
        fs = list(self.prev_f)[:order]
        dy = _lincomb(None, fs, [(1 / div) * c for c in coefs]) * dt
        if self.implicit:
            mc, mdiv = _MOULTON[order + 1]
            delta = _lincomb(None, fs, [(1 / mdiv) * c for c in mc[1:]]) * dt
            converged = False
            for _ in range(self.max_iters):
                dy_old = dy
                f = self.f(t + dt, y + dy.to(y.dtype))
                dy = dt * (mc[0] / mdiv) * f + delta
                converged = self._converged(dy_old, dy)
                if converged:
                    break
            if not converged:
                import sys
                print('Warning: Functional iteration did not converge. Solution may be incorrect.', file=sys.stderr)
                self.prev_f.pop()
            self._update_history(t, f)
        return y + dy.to(y.dtype)


class AdamsBashforth(AdamsBashforthMoulton):
    def __init__(self, func, y0, **kwargs):
        kwargs.pop('implicit', None)
        super(AdamsBashforth, self).__init__(func, y0, implicit=False, **kwargs)


SOLVERS = {
    'explicit_adams': AdamsBashforth,
    'fixed_adams': AdamsBashforthMoulton,
    'adams': VariableCoefficientAdamsBashforth,
    'tsit5': Tsit5Solver,
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
