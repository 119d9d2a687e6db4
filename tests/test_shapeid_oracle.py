"""Pins oracle/shapeid_oracle.py against fixtures produced by the reference's ShapeID package (CPU only)."""
import os

import numpy as np
import torch

from oracle import make_golden_shapeid as mgs
from oracle import shapeid_oracle as so

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "shapeid.npz"))


def test_perlin_shape_velocity_rhs_are_bit_exact():
    assert np.array_equal(so.perlin(mgs.SHAPE, mgs.RES, GOLD["lattice"]), GOLD["noise"])
    m, p = so.shape_from_noise(so.perlin(mgs.SHAPE, mgs.RES, GOLD["shape_lattice"]), mgs.PCT)
    assert np.array_equal(m, GOLD["shape_mask"]) and np.array_equal(p, GOLD["shape_prob"])
    pots = [so.perlin(mgs.SHAPE, mgs.RES, q) for q in GOLD["vel_lattices"]]
    V = so.curl_velocity(*pots, mgs.VMULT)
    for k, v in zip(("Vx", "Vy", "Vz"), V):
        assert np.array_equal(v, GOLD[k])
    assert np.array_equal(so.advect_rhs(GOLD["shape_prob"], V), GOLD["rhs0"])


def test_ode_solvers():
    V = [GOLD[k] for k in ("Vx", "Vy", "Vz")]
    t = list(np.arange(mgs.NT) * mgs.DT)
    for name, y0 in (("f64", GOLD["shape_prob"]), ("f32", GOLD["shape_prob"].astype(np.float32))):
        sol, trace, n_rhs = so.dopri5(lambda y: so.advect_rhs(y, V), y0, t, mgs.DT)
        ref = GOLD["dopri5_%s" % name]
        assert n_rhs == int(GOLD["dopri5_%s_nrhs" % name])
        assert np.abs(np.stack(sol) - ref).max() / np.abs(ref).max() < 2e-6
        for method in ("euler", "midpoint", "rk4"):
            sol = so.fixed(lambda y: so.advect_rhs(y, V), y0, t, method)
            assert np.array_equal(np.stack(sol), GOLD["%s_%s" % (method, name)])


def test_percentile_helper_matches_numpy():
    from brainfm_b200.ShapeID.perlin3d import _percentile_device
    rng = np.random.RandomState(0)
    for n in (10, 1001, 13440):
        x = rng.randn(n)
        for q in (85.0, 92.0, 99.9, 50.0, 0.0, 100.0, 33.3333):
            assert _percentile_device(torch.from_numpy(x), q) == np.percentile(x, q), (n, q)


def test_diffusion_rhs_oracle_matches_reference_fixture():
    """oracle.shapeid_oracle.diffuse_rhs against AdvDiffPDE.forward of the reference (tests/golden/pde.npz)."""
    import os
    import numpy as np
    from oracle import shapeid_oracle as so
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "pde.npz"))
    C32, C64 = G["C"].astype(np.float32), G["C"]
    for name, C, D, sp, neu in [("diff_const", C32, float(G["Dconst"]), (1, 1, 1), True),
                                ("diff_scalar", C32, G["D"], (1, 1, 1), True),
                                ("diff_scalar_nobc", C32, G["D"], (1, .8, 1.3), False),
                                ("diff_scalar_f64", C64, G["D"], (1, .8, 1.3), True)]:
        assert np.array_equal(so.diffuse_rhs(C, D, sp, neu), G["out_" + name]), name
