"""GPU parity of brainfm_b200.ShapeID against reference-generated fixtures and the numpy oracle.
Perlin noise (float64), shape mask, curl velocity, upwind RHS: bit-exact.  ODE solutions: 1e-5 relative to the
solution's magnitude, identical RHS-evaluation count and accept/reject trace."""
import os

import numpy as np
import pytest
import torch

from oracle import make_golden_shapeid as mgs
from oracle import shapeid_oracle as so

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "shapeid.npz"))


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_perlin_and_shape_bit_exact():
    from brainfm_b200.ShapeID import perlin3d as P
    np.random.seed(11)
    noise = P.generate_perlin_noise_3d(mgs.SHAPE, mgs.RES, tileable=(True, False, False))
    assert isinstance(noise, np.ndarray) and noise.dtype == np.float64
    assert np.array_equal(noise, GOLD["noise"])
    np.random.seed(12)
    mask, prob = P.generate_shape_3d(mgs.SHAPE, mgs.RES, mgs.PCT, 'cuda')
    assert mask.dtype == torch.float64 and prob.is_cuda
    assert np.array_equal(mask.cpu().numpy(), GOLD["shape_mask"])
    assert np.array_equal(prob.cpu().numpy(), GOLD["shape_prob"])
    np.random.seed(11)
    n2, m2 = P.generate_perlin_noise_3d(mgs.SHAPE, mgs.RES, tileable=(True, False, False), percentile=mgs.PCT)
    thr = np.percentile(GOLD["noise"], mgs.PCT)
    assert np.array_equal(m2, (GOLD["noise"] >= thr).astype(np.float64))
    with pytest.raises(ValueError):
        P.generate_perlin_noise_3d((25, 20, 28), mgs.RES)
    # fractal noise: octave sum of the same kernel
    np.random.seed(5)
    fr = P.generate_fractal_noise_3d((24, 24, 24), (2, 2, 2), octaves=2)
    np.random.seed(5)
    ref = so.perlin((24, 24, 24), (2, 2, 2), so.lattice((2, 2, 2), (False,) * 3)) * 1.0
    ref = ref + 0.5 * so.perlin((24, 24, 24), (4, 4, 4), so.lattice((4, 4, 4), (False,) * 3))
    assert np.array_equal(fr, ref)


def test_velocity_gradients_and_rhs_bit_exact():
    from brainfm_b200.ShapeID import perlin3d as P
    from brainfm_b200.ShapeID import misc as M
    from brainfm_b200.ShapeID.DiffEqs.pde import AdvDiffPDE
    np.random.seed(13)
    V = P.generate_velocity_3d(mgs.SHAPE, mgs.RES, mgs.VMULT, 'cuda')
    for k in ("Vx", "Vy", "Vz"):
        assert V[k].dtype == torch.float32
        assert np.array_equal(V[k].cpu().numpy(), GOLD[k]), k
    x = GOLD["noise"]
    gc = M.gradient_c(cu(x)).cpu().numpy()
    for ax, ref in enumerate(so.grad_c(x)):
        assert np.array_equal(gc[..., ax], ref)
    gf = M.gradient_f(cu(x.astype(np.float32))).cpu().numpy()
    assert np.array_equal(gf[:-1, :, :, 0], (x.astype(np.float32)[1:] - x.astype(np.float32)[:-1]))
    pde = AdvDiffPDE(data_spacing=[1., 1., 1.], perf_pattern='adv', V_type='vector_div_free', V_dict=V, BC='neumann',
                     dt=mgs.DT, device='cuda')
    out = pde(torch.tensor(0.), cu(GOLD["shape_prob"])[None])[0].cpu().numpy()
    assert out.dtype == np.float32 and np.array_equal(out, GOLD["rhs0"])
    with pytest.raises(NotImplementedError):       # tensor diffusivities are not built
        AdvDiffPDE([1., 1., 1.], 'adv_diff', D_type='full', V_type='vector_div_free', V_dict=V,
                   device='cuda')(0., cu(x)[None])


@pytest.mark.parametrize("name", ["f64", "f32"])
def test_ode_solvers_match_reference(name):
    from brainfm_b200.ShapeID.DiffEqs import odeint_adjoint, odeint
    from brainfm_b200.ShapeID.DiffEqs.pde import AdvDiffPDE
    V = {k: cu(GOLD[k]) for k in ("Vx", "Vy", "Vz")}
    pde = AdvDiffPDE(data_spacing=[1., 1., 1.], perf_pattern='adv', V_type='vector_div_free', V_dict=V, BC='neumann',
                     dt=mgs.DT, device='cuda')
    y0 = cu(GOLD["shape_prob"] if name == "f64" else GOLD["shape_prob"].astype(np.float32))
    t = torch.from_numpy(np.arange(mgs.NT) * mgs.DT).cuda()
    sol, solver = odeint_adjoint(pde, y0[None], t, mgs.DT, method='dopri5', return_solver=True)
    ref = GOLD["dopri5_%s" % name]
    assert sol.shape == (mgs.NT, 1, *mgs.SHAPE) and sol.dtype == y0.dtype
    assert solver.n_rhs == int(GOLD["dopri5_%s_nrhs" % name])          # integer parity of the control flow
    tr = GOLD["dopri5_%s_trace" % name]
    assert len(solver.trace) == len(tr)
    assert [bool(s[2]) for s in solver.trace] == [bool(r[2]) for r in tr]
    np.testing.assert_allclose([s[1] for s in solver.trace], tr[:, 1], rtol=1e-6)
    assert np.abs(sol[:, 0].cpu().numpy() - ref).max() / np.abs(ref).max() < 1e-5
    for method in ("euler", "midpoint", "rk4"):
        out = odeint(pde, y0[None], t, mgs.DT, method=method)[:, 0].cpu().numpy()
        ref = GOLD["%s_%s" % (method, name)]
        assert np.abs(out - ref).max() / np.abs(ref).max() < 1e-6, method
    with pytest.raises(KeyError):
        odeint(pde, y0[None], t, mgs.DT, method='bogus')
    with pytest.raises(ValueError):
        odeint_adjoint(lambda tt, y: y, y0[None], t, mgs.DT)


PDE = np.load(os.path.join(os.path.dirname(__file__), "golden", "pde.npz"))


@pytest.mark.parametrize("name,pattern,dtype_d,bc,spacing,dt,stoch", [
    ("diff_const", "diff", "constant", "neumann", [1., 1., 1.], "f32", False),
    ("diff_scalar", "diff", "scalar", "neumann", [1., 1., 1.], "f32", False),
    ("diff_scalar_nobc", "diff", "scalar", None, [1., 0.8, 1.3], "f32", False),
    ("diff_scalar_f64", "diff", "scalar", "neumann", [1., 0.8, 1.3], "f64", False),
    ("advdiff_scalar", "adv_diff", "scalar", "neumann", [1., 1., 1.], "f32", False),
    ("advdiff_const_f64", "adv_diff", "constant", None, [1., 1., 1.], "f64", False),
    ("advdiff_scalar_stoch0", "adv_diff", "scalar", "neumann", [1., 1., 1.], "f32", True)])
def test_diffusion_and_advection_diffusion_rhs_bit_exact(name, pattern, dtype_d, bc, spacing, dt, stoch):
    """AdvDiffPDE.forward for the 'diff' / 'adv_diff' patterns (bfm_diffuse_rhs) against the reference's output
    (tests/golden/pde.npz): separately rounded float32 differences, so the result is bit-exact."""
    from brainfm_b200.ShapeID.DiffEqs.pde import AdvDiffPDE
    C = cu(PDE["C"] if dt == "f64" else PDE["C"].astype(np.float32))[None]
    D = {"D": cu(PDE["D"])[None]} if dtype_d == "scalar" else {"D": float(PDE["Dconst"])}
    V = {k: cu(PDE[k]) for k in ("Vx", "Vy", "Vz")}
    pde = AdvDiffPDE(data_spacing=spacing, perf_pattern=pattern, D_type=dtype_d, V_type='vector_div_free', BC=bc,
                     dt=0.1, V_dict=V, D_dict=D, stochastic=stoch, device='cuda')
    out = pde(torch.tensor(0.), C)[0].cpu().numpy()
    assert out.dtype == np.float32 and np.array_equal(out, PDE["out_" + name])
    with pytest.raises(NotImplementedError):
        AdvDiffPDE(spacing, 'diff', D_type='full', V_type='vector_div_free', D_dict=D, device='cuda')(0., C)
