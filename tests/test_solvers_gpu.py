"""The remaining ODE methods of ShapeID/DiffEqs/odeint.py:8-17 -- tsit5, adams, fixed_adams, explicit_adams -- against
outputs of the reference itself (tests/golden/solvers.npz, oracle/make_golden_solvers.py) on a small 3-D advection
problem: solutions within 1e-5 relative / 1e-4 absolute of max|y|, step traces (time, step size, accept / reject,
order) and RHS-evaluation counts equal."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "solvers.npz"))
RTOL, ATOL, DT = float(GOLD["rtol"]), float(GOLD["atol"]), float(GOLD["dt"])


def _problem():
    from brainfm_b200.ShapeID.DiffEqs.pde import AdvDiffPDE
    dev = torch.device("cuda", 0)
    pde = AdvDiffPDE(data_spacing=[1., 1., 1.], perf_pattern='adv', V_type='vector_div_free', V_dict={}, BC='neumann',
                     dt=DT, device=dev)
    pde.V_dict = {k: torch.from_numpy(GOLD[k]).to(dev) for k in ("Vx", "Vy", "Vz")}
    y0 = torch.from_numpy(GOLD["y0"]).to(dev)
    return pde, y0


def _close(got, want, what):
    scale = float(np.abs(want).max())
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-4 * max(scale, 1.0), err_msg=what)


def test_adams_matches_the_reference():
    from brainfm_b200.ShapeID.DiffEqs.odeint import odeint
    pde, y0 = _problem()
    t = torch.from_numpy(np.arange(4) * DT)
    with torch.no_grad():
        sol, solver = odeint(pde, y0[None], t, DT, method='adams', rtol=RTOL, atol=ATOL, return_solver=True)
    ref_trace = GOLD["adams_trace"]
    got = np.array([[a, b, float(c), o] for a, b, c, r, o in solver.trace])
    assert got.shape == ref_trace.shape, (got.shape, ref_trace.shape)
    assert np.array_equal(got[:, 2:], ref_trace[:, 2:])                     # accept / reject and order, step by step
    np.testing.assert_allclose(got[:, :2], ref_trace[:, :2], rtol=1e-4, atol=1e-9)
    assert solver.n_rhs == int(GOLD["adams_nrhs"])
    _close(sol[:, 0].cpu().numpy(), GOLD["adams"], "adams solution")


def test_tsit5_reproduces_the_reference_trace():
    """The first 40 step attempts: start time, tried step size and accept / reject decision."""
    from brainfm_b200.ShapeID.DiffEqs.odeint import Tsit5Solver
    pde, y0 = _problem()
    solver = Tsit5Solver(pde, y0[None], rtol=RTOL, atol=ATOL, dt=DT, max_num_steps=40)
    with torch.no_grad(), pytest.raises(AssertionError, match="max_num_steps exceeded"):
        solver.integrate(torch.tensor([0.0, 1.0], dtype=torch.float64))
    ref = GOLD["tsit5_trace"]                                     # rows: state.t0, state.t1, dt tried, next dt
    assert len(solver.trace) == 40
    t_start = np.array([r[0] for r in solver.trace])
    dt_try = np.array([r[1] for r in solver.trace])
    accepted = np.array([r[2] for r in solver.trace])
    np.testing.assert_allclose(t_start, ref[:, 0], rtol=1e-4, atol=1e-12)
    np.testing.assert_allclose(dt_try, ref[:, 2], rtol=1e-4)
    assert np.array_equal(accepted, ref[:, 1] > ref[:, 0])


def test_tsit5_short_solve_with_dense_output():
    from brainfm_b200.ShapeID.DiffEqs.odeint import odeint
    pde, y0 = _problem()
    t = torch.from_numpy(GOLD["tsit5_t"])
    with torch.no_grad():
        sol, solver = odeint(pde, y0[None], t, DT, method='tsit5', rtol=RTOL, atol=ATOL, return_solver=True)
    assert solver.n_rhs == int(GOLD["tsit5_nrhs"])
    _close(sol[:, 0].cpu().numpy(), GOLD["tsit5_sol"], "tsit5 dense output")


@pytest.mark.parametrize("method", ["fixed_adams", "explicit_adams"])
def test_fixed_grid_adams(method):
    """The reference with its missing `rk_common` name supplied (it raises NameError as shipped -- recorded in the
    fixture): RK4 bootstrap, then Adams-Bashforth(-Moulton) with the corrector iteration."""
    from brainfm_b200.ShapeID.DiffEqs.odeint import odeint
    assert int(GOLD[method + "_reference_raises"]) == 1
    pde, y0 = _problem()
    t = torch.from_numpy(np.arange(9) * 0.02)
    with torch.no_grad():
        sol, solver = odeint(pde, y0[None], t, DT, method=method, rtol=RTOL, atol=ATOL, return_solver=True)
    assert solver.n_rhs == int(GOLD[method + "_nrhs"])
    _close(sol[:, 0].cpu().numpy(), GOLD[method], method)
