"""TEST INFRASTRUCTURE ONLY.  Runs the reference's ShapeID (perlin3d, stream_3D, AdvDiffPDE, odeint_adjoint) on a
small seeded problem, stores tests/golden/shapeid.npz and checks oracle/shapeid_oracle.py against it.

    python -m oracle.make_golden_shapeid     (build container only)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim as rs           # noqa: E402
from oracle import shapeid_oracle as so     # noqa: E402

# V_multiplier 60 on a 24^3 grid gives |V| ~ 12 voxels per unit time, the regime of the 192^3 configuration
# (500 there); with 500 on this small grid the forced 0.02 steps are CFL-unstable and values reach 1e12.
SHAPE, RES, PCT, VMULT, DT, NT = (24, 20, 28), [2, 2, 2], 92.0, 60, 0.1, 4


def main():
    rs.install()
    from ShapeID.perlin3d import generate_shape_3d, generate_velocity_3d, generate_perlin_noise_3d
    from ShapeID.DiffEqs.adjoint import odeint_adjoint
    from ShapeID.DiffEqs.pde import AdvDiffPDE
    gold = {}
    np.random.seed(11)
    noise = generate_perlin_noise_3d(SHAPE, RES, tileable=(True, False, False))
    np.random.seed(11)
    g = so.lattice(RES)
    assert np.array_equal(so.perlin(SHAPE, RES, g), noise), "perlin oracle is not bit-exact"
    gold["lattice"], gold["noise"] = g, noise
    np.random.seed(12)
    mask, prob = generate_shape_3d(SHAPE, RES, PCT, 'cpu')
    np.random.seed(12)
    g2 = so.lattice(RES)
    m2, p2 = so.shape_from_noise(so.perlin(SHAPE, RES, g2), PCT)
    assert np.array_equal(m2, mask.numpy()) and np.array_equal(p2, prob.numpy())
    gold["shape_lattice"], gold["shape_mask"], gold["shape_prob"] = g2, mask.numpy(), prob.numpy()
    np.random.seed(13)
    V = generate_velocity_3d(SHAPE, RES, VMULT, 'cpu')
    np.random.seed(13)
    gl = [so.lattice(RES) for _ in range(3)]
    pots = [so.perlin(SHAPE, RES, q) for q in gl]
    Vo = so.curl_velocity(*pots, VMULT)
    for k, v in zip(("Vx", "Vy", "Vz"), Vo):
        assert np.array_equal(v, V[k].numpy()), k
        gold[k] = V[k].numpy()
    gold["vel_lattices"] = np.stack(gl)
    pde = AdvDiffPDE(data_spacing=[1., 1., 1.], perf_pattern='adv', V_type='vector_div_free', V_dict={}, BC='neumann',
                     dt=DT, device='cpu')
    pde.V_dict = V
    rhs_ref = pde(torch.tensor(0.), prob[None])[0].numpy()
    Vn = [V[k].numpy() for k in ("Vx", "Vy", "Vz")]
    assert np.array_equal(so.advect_rhs(prob.numpy(), Vn), rhs_ref), "advect oracle is not bit-exact"
    gold["rhs0"] = rhs_ref
    t = torch.from_numpy(np.arange(NT) * DT)
    count = [0]
    orig = pde.forward

    def counted(tt, c):
        count[0] += 1
        return orig(tt, c)
    pde.forward = counted
    for state, name in ((prob, "f64"), (prob.float(), "f32")):
        count[0] = 0
        with torch.no_grad():
            ref = odeint_adjoint(pde, state[None], t, DT, method='dopri5')[:, 0].numpy()
        sol, trace, n_rhs = so.dopri5(lambda y: so.advect_rhs(y, Vn), state.numpy(), list(np.arange(NT) * DT), DT)
        d = float(np.abs(np.stack(sol) - ref).max() / np.abs(ref).max())
        print("dopri5 %s: reference RHS evals %d, oracle %d, steps %d, max|diff|/max|ref| %.3e (max|ref| %.3e)" % (
            name, count[0], n_rhs, len(trace), d, np.abs(ref).max()))
        assert n_rhs == count[0] and d < 2e-6
        gold["dopri5_%s" % name] = ref
        gold["dopri5_%s_nrhs" % name] = np.array(count[0])
        gold["dopri5_%s_trace" % name] = np.array([[a, b, float(c), r] for a, b, c, r in trace])
        for method in ("euler", "midpoint", "rk4"):
            with torch.no_grad():
                ref = odeint_adjoint(pde, state[None], t, DT, method=method)[:, 0].numpy()
            sol = so.fixed(lambda y: so.advect_rhs(y, Vn), state.numpy(), list(np.arange(NT) * DT), method)
            d = float(np.abs(np.stack(sol) - ref).max() / np.abs(ref).max())
            print("  %s %s max|diff|/max|ref| %.3e" % (method, name, d))
            assert d < 2e-6
            gold["%s_%s" % (method, name)] = ref
    gold["meta.versions"] = np.array("torch %s numpy %s" % (torch.__version__, np.__version__))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "shapeid.npz"), **gold)
    print("shapeid fixtures:", len(gold), "arrays")


if __name__ == "__main__":
    main()
