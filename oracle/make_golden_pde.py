"""TEST INFRASTRUCTURE ONLY.  AdvDiffPDE.forward of the unmodified reference (ShapeID/DiffEqs/pde.py:563-640) for the
diffusion and advection-diffusion patterns, constant and scalar diffusivity, with and without the Neumann boundary,
float32 and float64 states, unit and anisotropic spacing.  -> tests/golden/pde.npz
    python -m oracle.make_golden_pde     (build container only)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim as rs           # noqa: E402

SHAPE = (12, 10, 14)
CASES = [  # name, perf_pattern, D_type, BC, spacing, dtype, stochastic
    ("diff_const", "diff", "constant", "neumann", [1., 1., 1.], "f32", False),
    ("diff_scalar", "diff", "scalar", "neumann", [1., 1., 1.], "f32", False),
    ("diff_scalar_nobc", "diff", "scalar", None, [1., 0.8, 1.3], "f32", False),
    ("diff_scalar_f64", "diff", "scalar", "neumann", [1., 0.8, 1.3], "f64", False),
    ("advdiff_scalar", "adv_diff", "scalar", "neumann", [1., 1., 1.], "f32", False),
    ("advdiff_const_f64", "adv_diff", "constant", None, [1., 1., 1.], "f64", False),
    ("advdiff_scalar_stoch0", "adv_diff", "scalar", "neumann", [1., 1., 1.], "f32", True),
]


def main():
    rs.install()
    from ShapeID.DiffEqs.pde import AdvDiffPDE
    rng = np.random.RandomState(5)
    gold = {"C": rng.rand(*SHAPE), "D": (0.1 + rng.rand(*SHAPE)).astype(np.float32), "Dconst": np.array(0.37)}
    for k in ("Vx", "Vy", "Vz"):
        gold[k] = (4 * rng.randn(*SHAPE)).astype(np.float32)
    for name, pattern, dtype_d, bc, spacing, dt, stoch in CASES:
        C = torch.from_numpy(gold["C"] if dt == "f64" else gold["C"].astype(np.float32))[None]
        D = {"D": torch.from_numpy(gold["D"])[None]} if dtype_d == "scalar" else {"D": torch.tensor(float(gold["Dconst"]))}
        V = {k: torch.from_numpy(gold[k])[None] for k in ("Vx", "Vy", "Vz")}
        pde = AdvDiffPDE(data_spacing=spacing, perf_pattern=pattern, D_type=dtype_d, V_type='vector_div_free', BC=bc,
                         dt=0.1, V_dict=V, D_dict=D, stochastic=stoch, device='cpu')
        out = pde(torch.tensor(0.), C)
        gold["out_" + name] = out[0].numpy()
        print("%-24s %s max|out| %.4g" % (name, out.dtype, float(out.abs().max())))
    gold["meta.versions"] = np.array("torch %s numpy %s" % (torch.__version__, np.__version__))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pde.npz"), **gold)


if __name__ == "__main__":
    main()
