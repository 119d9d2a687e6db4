"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's ShapeID path.  Never imported by the product.

  perlin()            ShapeID/perlin3d.py:15-90     (float64, same operation order => bit-exact)
  curl_velocity()     ShapeID/misc.py:66-80,198-259 + perlin3d.py:149-156
  advect_rhs()        ShapeID/DiffEqs/pde.py:588-640, 301-328, 499-509 (Neumann BC, upwind)
  dopri5()/fixed()    ShapeID/DiffEqs/dopri5.py:58-175, rk_common.py:22-80, interp.py:5-65, misc.py:84-170,
                      fixed_grid.py, solvers.py:103-216
Pinned against the reference by oracle/make_golden_shapeid.py -> tests/golden/shapeid.npz."""
import math

import numpy as np

F32, F64 = np.float32, np.float64


def lattice(res, tileable=(True, False, False)):
    theta = 2 * np.pi * np.random.rand(res[0] + 1, res[1] + 1, res[2] + 1)
    phi = 2 * np.pi * np.random.rand(res[0] + 1, res[1] + 1, res[2] + 1)
    g = np.stack((np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta), np.cos(phi)), axis=3)
    if tileable[0]:
        g[-1] = g[0]
    if tileable[1]:
        g[:, -1] = g[:, 0]
    if tileable[2]:
        g[:, :, -1] = g[:, :, 0]
    return g


def perlin(shape, res, grad):
    """Noise from a given gradient lattice (res+1)^3 x 3."""
    ax = []
    for s, r in zip(shape, res):
        i = np.arange(s)
        ax.append((np.mod(i * (r / s), 1.0), i // (s // r)))
    (gx, a), (gy, b), (gz, c) = ax
    GX, GY, GZ = np.meshgrid(gx, gy, gz, indexing="ij")
    A, B, Cc = np.meshgrid(a, b, c, indexing="ij")

    def ramp(da, db, dc):
        g = grad[A + da, B + db, Cc + dc]
        return ((GX - da) * g[..., 0] + (GY - db) * g[..., 1]) + (GZ - dc) * g[..., 2]

    def fade(t):
        return t * t * t * (t * (t * 6 - 15) + 10)

    t0, t1, t2 = fade(GX), fade(GY), fade(GZ)
    n00 = ramp(0, 0, 0) * (1 - t0) + t0 * ramp(1, 0, 0)
    n10 = ramp(0, 1, 0) * (1 - t0) + t0 * ramp(1, 1, 0)
    n01 = ramp(0, 0, 1) * (1 - t0) + t0 * ramp(1, 0, 1)
    n11 = ramp(0, 1, 1) * (1 - t0) + t0 * ramp(1, 1, 1)
    n0 = (1 - t1) * n00 + t1 * n10
    n1 = (1 - t1) * n01 + t1 * n11
    return (1 - t2) * n0 + t2 * n1


def shape_from_noise(noise, percentile):
    thr = np.percentile(noise, percentile)
    mask = (noise >= thr).astype(F64)
    return mask, noise * mask


def grad_c(X):
    """gradient_c components as float32 (central, one-sided at the borders)."""
    out = []
    for ax in range(3):
        Xm = np.moveaxis(X, ax, 0)
        d = np.zeros(Xm.shape, dtype=F32)
        d[1:-1] = ((Xm[2:] - Xm[:-2]) / 2).astype(F32)
        d[0] = (Xm[1] - Xm[0]).astype(F32)
        d[-1] = (Xm[-1] - Xm[-2]).astype(F32)
        out.append(np.moveaxis(d, 0, ax))
    return out


def curl_velocity(a, b, c, mult):
    da, db, dc = grad_c(a), grad_c(b), grad_c(c)
    m = F32(mult)
    return (dc[1] - db[2]) * m, (da[2] - dc[0]) * m, (db[0] - da[1]) * m


def advect_rhs(C, V, neumann=True):
    """C: (D,H,W) float32/float64; V: three float32 volumes -> float32."""
    if neumann:
        C = np.pad(C[1:-1, 1:-1, 1:-1], 1, mode="edge")
    terms = []
    for ax in range(3):
        Cm = np.moveaxis(C, ax, 0)
        df = np.zeros(Cm.shape, dtype=F32)
        db = np.zeros(Cm.shape, dtype=F32)
        df[:-1] = (Cm[1:] - Cm[:-1]).astype(F32)
        df[-1] = (Cm[-1] - Cm[-2]).astype(F32)
        db[1:] = (Cm[1:] - Cm[:-1]).astype(F32)
        db[0] = (Cm[1] - Cm[0]).astype(F32)
        df, db = np.moveaxis(df, 0, ax), np.moveaxis(db, 0, ax)
        terms.append(V[ax] * np.where(V[ax] > 0, db, df))
    return -((terms[0] + terms[1]) + terms[2])


ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1., 1.]
BETA = [[1 / 5], [3 / 40, 9 / 40], [44 / 45, -56 / 15, 32 / 9],
        [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
        [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
        [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]]
C_ERROR = [35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
           -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1. / 60.]
C_MID = [6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
         187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2]


def _sdp(dt, coefs, ks, T):
    acc = None
    for c, k in zip(coefs, ks):
        term = F32(T(dt) * c) * k
        acc = term if acc is None else acc + term
    return acc


def _rms(x):
    return float(np.sqrt(np.sum(x.astype(F64) ** 2)) / math.sqrt(x.size))


def dopri5(rhs, y0, t, dt_cfg, rtol=1e-6, atol=1e-12):
    """Returns (solutions list, trace [(t0, dt, accepted, ratio)], n_rhs)."""
    T = y0.dtype.type
    n_rhs = [0]

    def f(y):
        n_rhs[0] += 1
        return rhs(y)

    f0 = f(y0)
    scale = atol + np.abs(y0) * rtol
    d0, d1 = _rms(y0 / scale), _rms(f0 / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * (d0 / d1)
    h0 = float(T(h0))
    f1 = f(y0 + F32(h0) * f0)
    d2 = _rms((f1 - f0) / scale) / h0
    h1 = max(1e-6, h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** (1. / 5.)
    dt = min(100 * h0, h1)
    t1 = t[0]
    sol, trace, last = [y0], [], None
    tol_min = 0.2 * dt_cfg if 0.1 * dt_cfg >= 0.01 else 0.01
    for nt in t[1:]:
        while nt > t1:
            ks = [f0]
            yi = y0
            for beta in BETA:
                yi = y0 + _sdp(dt, beta, ks, T).astype(y0.dtype)
                ks.append(f(yi))
            y1 = yi
            err = _sdp(dt, C_ERROR, ks, T)
            tol = T(atol) + T(rtol) * np.maximum(np.abs(y0), np.abs(y1))
            r = err.astype(y0.dtype) / tol
            ratio = float(np.mean((r * r).astype(F64)))
            accept = ratio <= 1
            if ratio == 0:
                dt_next = dt * 10.0
            else:
                dfac = 1.0 if ratio < 1 else 0.2
                dt_next = dt / max(0.1, min((ratio ** 0.5) ** 0.2 / 0.9, 1 / dfac))
            forced = dt_next < tol_min or dt_next > 0.1
            if forced:
                dt_next = tol_min if dt_next < tol_min else dt_next
                dt_next = 0.1 if dt_next > 0.1 else dt_next
            trace.append((t1, dt, bool(accept or forced), ratio))
            if accept or forced:
                last = (y0, y1, ks, dt, t1, t1 + dt)
                t1 = t1 + dt
                y0, f0 = y1, ks[-1]
            dt = dt_next
        a0, a1, ks, h, ta, tb = last
        hh = float(T(h))
        ym = a0 + _sdp(h, C_MID, ks, T).astype(a0.dtype)
        fa, fb = ks[0], ks[-1]
        ca = F32(-2 * hh) * fa + F32(2 * hh) * fb + -8 * a0 + -8 * a1 + 16 * ym
        cb = F32(5 * hh) * fa + F32(-3 * hh) * fb + 18 * a0 + 14 * a1 + -32 * ym
        cc = F32(-4 * hh) * fa + F32(hh) * fb + -11 * a0 + -5 * a1 + 16 * ym
        cd = F32(hh) * fa
        x = T((T(nt) - T(ta)) / (T(tb) - T(ta)))
        x2 = x * x
        x3 = x2 * x
        x4 = x3 * x
        sol.append(ca * x4 + cb * x3 + cc * x2 + cd * F32(x) + a0 * T(1))
    return sol, trace, n_rhs[0]


def fixed(rhs, y0, t, method):
    T = y0.dtype.type
    y = y0
    sol = [y]
    tt = [T(v) for v in t]
    for t0, t1 in zip(tt[:-1], tt[1:]):
        dt = float(T(t1 - t0))
        d32 = F32(dt)
        if method == "euler":
            y = y + (d32 * rhs(y)).astype(y.dtype)
        elif method == "midpoint":
            ym = y + (rhs(y) * d32 / 2)
            y = y + (d32 * rhs(ym.astype(y.dtype))).astype(y.dtype)
        else:
            k1 = rhs(y)
            k2 = rhs((y + d32 * k1 / 3).astype(y.dtype))
            k3 = rhs((y + d32 * (k1 / -3 + k2)).astype(y.dtype))
            k4 = rhs((y + d32 * (k1 - k2 + k3)).astype(y.dtype))
            y = (y + (k1 + 3 * k2 + 3 * k3 + k4) * F32(dt / 8)).astype(y.dtype)
        sol.append(y)
    return sol


# ---- diffusion right-hand side (ShapeID/DiffEqs/pde.py:13-183, 331-353, 551-559, 623-639) ----------------------
def _set_bc(C):
    """set_BC for 'neumann' / 'cauchy': replicate-pad the interior (pde.py:590-600)."""
    return np.pad(C[1:-1, 1:-1, 1:-1], 1, mode="edge")


def _grad(X, mode, spacing):
    """gradient_f / gradient_b / gradient_c of a 3-D array: float32 buffers divided by the spacing."""
    out = []
    for d in range(3):
        Xd = np.moveaxis(X, d, 0)
        g = np.zeros(Xd.shape, dtype=np.float32)
        if mode == "f":
            g[:-1] = Xd[1:] - Xd[:-1]
            g[-1] = Xd[-1] - Xd[-2]
        elif mode == "b":
            g[1:] = Xd[1:] - Xd[:-1]
            g[0] = Xd[1] - Xd[0]
        else:
            g[1:-1] = (Xd[2:] - Xd[:-2]) / 2
            g[0] = Xd[1] - Xd[0]
            g[-1] = Xd[-1] - Xd[-2]
        out.append(np.moveaxis(g / np.float32(spacing[d]), 0, d))
    return out


def diffuse_rhs(C, D, spacing=(1., 1., 1.), neumann=True):
    """Grad_constantD (D a scalar) / Grad_scalarD (D an array) of the reference, same composition of differences."""
    C = _set_bc(C) if neumann else C
    gf = _grad(C, "f", spacing)
    dd = [_grad(gf[d], "b", spacing)[d] for d in range(3)]
    if np.ndim(D) == 0:
        return (np.float32(D) * ((dd[0] + dd[1]) + dd[2])).astype(np.float32)
    D = np.asarray(D, dtype=np.float32)
    gD, gC = _grad(D, "c", spacing), _grad(C, "c", spacing)
    out = (gD[0] * gC[0] + gD[1] * gC[1]) + gD[2] * gC[2]
    for d in range(3):
        out = out + D * dd[d]
    return out.astype(np.float32)
