"""TEST INFRASTRUCTURE ONLY.  Pins the remaining ODE methods of ShapeID/DiffEqs/odeint.py:8-17 (tsit5, adams,
fixed_adams, explicit_adams) on outputs of the UNMODIFIED reference, on the small advection problem of
make_golden_shapeid.py (24 x 20 x 28 grid, curl velocity, float32 state).  -> tests/golden/solvers.npz

    python -m oracle.make_golden_solvers     (build container only)

* adams: full solve, + per-step (t_n, dt, accepted, order) trace captured by wrapping _adaptive_adams_step.
* tsit5: the reference's error estimate rejects steps until dt ~ 1e-7 (a 2-element linear ODE on [0, 1] takes 300 s),
  so the fixture is the (t0, t1, dt) trace of the first 40 calls of _adaptive_tsit5_step, plus the dense output the
  reference would return after them.
* fixed_adams / explicit_adams: the reference raises NameError on the first step (fixed_adams.py:165 uses `rk_common`,
  which `import ShapeID.DiffEqs.rk_common` at :5 never binds); the fixture is produced with that ONE module attribute
  supplied (`fixed_adams.rk_common = ShapeID.DiffEqs.rk_common`) and records that it was needed.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim as rs           # noqa: E402

SHAPE, RES, PCT, VMULT, DT = (24, 20, 28), [2, 2, 2], 92.0, 60, 0.1
RTOL, ATOL = 1e-3, 1e-5


def problem():
    from ShapeID.perlin3d import generate_shape_3d, generate_velocity_3d
    from ShapeID.DiffEqs.pde import AdvDiffPDE
    np.random.seed(12)
    mask, prob = generate_shape_3d(SHAPE, RES, PCT, 'cpu')
    np.random.seed(13)
    V = generate_velocity_3d(SHAPE, RES, VMULT, 'cpu')
    pde = AdvDiffPDE(data_spacing=[1., 1., 1.], perf_pattern='adv', V_type='vector_div_free', V_dict={}, BC='neumann',
                     dt=DT, device='cpu')
    pde.V_dict = V
    return pde, prob.float(), V


def main():
    rs.install()
    import ShapeID.DiffEqs.rk_common
    from ShapeID.DiffEqs import adams, fixed_adams, tsit5
    from ShapeID.DiffEqs.misc import _check_inputs
    from ShapeID.DiffEqs.odeint import odeint
    gold = {"rtol": np.array(RTOL), "atol": np.array(ATOL), "dt": np.array(DT)}
    pde, y0, V = problem()
    gold["y0"] = y0.numpy()
    for k in ("Vx", "Vy", "Vz"):
        gold[k] = V[k].numpy()
    count = [0]
    orig = pde.forward

    def counted(tt, c):
        count[0] += 1
        return orig(tt, c)
    pde.forward = counted

    # ---- adams
    t = torch.from_numpy(np.arange(4) * DT)
    trace = []
    inner = adams.VariableCoefficientAdamsBashforth._adaptive_adams_step

    def traced(self, state, final_t):
        t_n, order = float(state.prev_t[0]), float(state.order)      # prev_t is mutated in place by an accepted step
        nt = min(float(state.next_t), float(final_t))
        new = inner(self, state, final_t)
        trace.append([t_n, nt - t_n, float(new.y_n is not state.y_n), order])
        return new
    adams.VariableCoefficientAdamsBashforth._adaptive_adams_step = traced
    count[0] = 0
    with torch.no_grad():
        sol = odeint(pde, y0[None], t, DT, method='adams', rtol=RTOL, atol=ATOL)
    adams.VariableCoefficientAdamsBashforth._adaptive_adams_step = inner
    gold["adams"] = sol[:, 0].numpy()
    gold["adams_trace"] = np.array(trace)
    gold["adams_nrhs"] = np.array(count[0])
    print("adams: %d steps (%d accepted), %d RHS evaluations, max order %d, max|y| %.3e" % (
        len(trace), int(sum(r[2] for r in trace)), count[0], int(max(r[3] for r in trace)), float(sol.abs().max())))

    # ---- tsit5: first 40 step attempts
    _, func, y0t, tt = _check_inputs(pde, y0[None], t)
    s = tsit5.Tsit5Solver(func, y0t, rtol=RTOL, atol=ATOL)
    with torch.no_grad():
        s.before_integrate(tt)
        rows = []
        for _ in range(40):
            dt_try = float(s.rk_state.dt)
            s.rk_state = s._adaptive_tsit5_step(s.rk_state)
            st = s.rk_state
            rows.append([float(st.t0), float(st.t1), dt_try, float(st.dt)])
        t_mid = 0.5 * (float(st.t0) + float(st.t1)) if float(st.t1) > float(st.t0) else float(st.t1)
        dense = tsit5._interp_eval_tsit5(st.t0, st.t1, st.interp_coeff, torch.tensor(t_mid, dtype=torch.float64)) \
            if float(st.t1) > float(st.t0) else None
    gold["tsit5_trace"] = np.array(rows)
    gold["tsit5_y1"] = st.y1[0][0].numpy()
    if dense is not None:
        gold["tsit5_dense_t"] = np.array(t_mid)
        gold["tsit5_dense"] = dense[0][0].numpy()
    print("tsit5: 40 attempts, %d accepted, t1 = %.3e, last dt %.3e" % (
        sum(1 for a, b in zip(rows[:-1], rows[1:]) if b[1] > a[1]) + (rows[0][1] > 0), rows[-1][1], rows[-1][3]))

    # ... and a (short) full solve through odeint, dense output included
    t_short = torch.tensor([0.0, 1e-3, 2e-3], dtype=torch.float64)
    count[0] = 0
    with torch.no_grad():
        sol = odeint(pde, y0[None], t_short, DT, method='tsit5', rtol=RTOL, atol=ATOL)
    gold["tsit5_t"], gold["tsit5_sol"], gold["tsit5_nrhs"] = t_short.numpy(), sol[:, 0].numpy(), np.array(count[0])
    print("tsit5 solve to t = 2e-3: %d RHS evaluations, max|y| %.3e" % (count[0], float(sol.abs().max())))

    # ---- fixed_adams / explicit_adams (reference + the one missing module attribute)
    t8 = torch.from_numpy(np.arange(9) * 0.02)
    for method in ("fixed_adams", "explicit_adams"):
        try:
            odeint(pde, y0[None], t8, DT, method=method, rtol=RTOL, atol=ATOL)
            gold[method + "_reference_raises"] = np.array(0)
        except NameError as e:
            gold[method + "_reference_raises"] = np.array(1)
            print("%s: the unmodified reference raises NameError(%s)" % (method, e))
        fixed_adams.rk_common = ShapeID.DiffEqs.rk_common
        count[0] = 0
        with torch.no_grad():
            sol = odeint(pde, y0[None], t8, DT, method=method, rtol=RTOL, atol=ATOL)
        del fixed_adams.rk_common
        gold[method] = sol[:, 0].numpy()
        gold[method + "_nrhs"] = np.array(count[0])
        print("%s: %d RHS evaluations over 8 steps, max|y| %.3e" % (method, count[0], float(sol.abs().max())))
    gold["meta.versions"] = np.array("torch %s numpy %s" % (torch.__version__, np.__version__))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "solvers.npz"), **gold)
    print("solver fixtures:", len(gold), "arrays")


if __name__ == "__main__":
    main()
