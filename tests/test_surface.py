"""read_and_deform_surface (Generator/utils.py:479-533): oracle vs the reference-generated fixture (CPU), CUDA path vs
both (GPU).  Vertices within 1e-5 rel / 1e-4 abs, faces (integers) exact, flip swaps left and right."""
import os

import numpy as np
import pytest
import torch

from oracle import gen_oracle as go
from oracle import make_golden_surface as mgs

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "surface.npz"))
KEYS = ('Vlw', 'Flw', 'Vrw', 'Frw', 'Vlp', 'Flp', 'Vrp', 'Frp')


@pytest.mark.parametrize("name", sorted(mgs.CASES))
def test_oracle_matches_reference_fixture(name):
    seed, flip = mgs.CASES[name]
    mat, A, c2, Fneg = mgs.inputs(seed)
    out = go.surface_deform(mat, A, c2, Fneg, flip, list(mgs.SIZE))
    for k in KEYS:
        ref = GOLD["%s_%s" % (name, k)]
        if k[0] == 'F':
            assert np.array_equal(out[k].numpy(), ref), k
        else:
            assert np.array_equal(out[k].numpy(), ref), "%s: oracle must reproduce the reference bit for bit" % k


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mgs.CASES))
def test_cuda_surface_matches_fixture_and_oracle(name):
    from brainfm_b200 import io as bio
    from brainfm_b200.Generator.constants import processing_funcs
    seed, flip = mgs.CASES[name]
    mat, A, c2, Fneg = mgs.inputs(seed)
    path = "/virtual/%s.nii" % name
    bio.register_surface("/virtual/%s.mat" % name, mat)
    dd = {"Fneg": Fneg.cuda(), "A": A.cuda(), "c2": c2.cuda()}
    got = processing_funcs['surface'](None, 'surface', path, {"flip": flip}, dd, 'cuda', None, list(mgs.SIZE))
    assert list(got) == list(KEYS)
    orc = go.surface_deform(mat, A, c2, Fneg, flip, list(mgs.SIZE))
    for k in KEYS:
        ref = GOLD["%s_%s" % (name, k)]
        g = got[k].cpu().numpy()
        if k[0] == 'F':
            assert got[k].dtype == torch.int32 and np.array_equal(g, ref), k
        else:
            np.testing.assert_allclose(g, ref, rtol=1e-5, atol=1e-4, err_msg=k)
            np.testing.assert_allclose(g, orc[k].numpy(), rtol=1e-5, atol=1e-4, err_msg=k)
    with pytest.raises(ValueError):
        processing_funcs['surface'](None, 'surface', path, {"flip": flip}, {"Fneg": None, "A": A, "c2": c2}, 'cuda')
