"""Pins oracle/interpol_oracle.py against fixtures produced by the reference's utils/interpol (CPU only)."""
import os

import numpy as np
import pytest

from oracle import interpol_oracle as io
from oracle import make_golden_interpol as mgi

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "interpol.npz"))
CASES = [c for c in mgi.cases() if max(c[3]) <= 3 and c[0] in GOLD.files]


@pytest.mark.parametrize("chunk", range(8))
def test_oracle_pull_push_count_grad(chunk):
    n = 0
    for key, kind, dim, order, bound, e in CASES[chunk::8]:
        vol, grid, push_in, ishape, oshape = mgi.problem(dim, 100 + dim)
        if kind == 'pull':
            out = io.pull(vol, grid, order, bound, e)
        elif kind == 'grad':
            out = io.pull(vol, grid, order, bound, e, grad=True)
        elif kind == 'push':
            out = io.push(push_in, grid, ishape, order, bound, e)
        else:
            out = io.push(None, grid, ishape, order, bound, e)
        np.testing.assert_allclose(out, GOLD[key], rtol=1e-5, atol=2e-5, err_msg=key)
        n += 1
    assert n > 50


def test_oracle_prefilter():
    for n in (5, 40):
        x = GOLD["coeff_in_n%d" % n]
        for o in (2, 3):
            for b in (0, 1, 2, 3, 6):
                np.testing.assert_allclose(io.spline_filter(x, o, b, 1), GOLD["coeff_n%d_o%d_b%d" % (n, o, b)],
                                           rtol=1e-5, atol=2e-5)


def test_appendix_c_known_answers():
    """SURVEY.md appendix C (reference-generated 1-D known answers)."""
    x = np.array([1., 2., 3., 4., 5.], dtype=np.float32)[None, None]
    g = np.arange(-4., 9., dtype=np.float32)[None, :, None]
    want = {"zero": [0, 0, 0, 0, 1, 2, 3, 4, 5, 0, 0, 0, 0], "dct2": [4, 3, 2, 1, 1, 2, 3, 4, 5, 5, 4, 3, 2],
            "dst1": [-3, -2, -1, 0, 0, 2, 3, 4, 5, 0, -5, -4, -3], "dft": [2, 3, 4, 5, 1, 2, 3, 4, 5, 1, 2, 3, 4]}
    for name, row in want.items():
        out = io.pull(x, g, [1], [io.BOUNDS[name]], 1)[0, 0]
        np.testing.assert_allclose(out, row, atol=1e-6)
        np.testing.assert_allclose(out, GOLD["appC_%s_o1" % name], atol=1e-6)
    for name in io.BOUNDS:
        for o in (0, 1, 3):
            np.testing.assert_allclose(io.pull(x, g, [o], [io.BOUNDS[name]], 1)[0, 0], GOLD["appC_%s_o%d" % (name, o)],
                                       atol=1e-5)
