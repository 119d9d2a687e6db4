"""utils/misc.py resamplers (myzoom_torch_anisotropic, torch_resize) against fixtures produced by the reference
(tests/golden/misc.npz, oracle/make_golden_misc.py): zoom outputs are pure separately-rounded lerps (bit-exact), the
blurred ones within the float tolerance; affines equal."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "misc.npz"))


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("name", ["up", "down", "mixed"])
def test_myzoom_torch_anisotropic(name):
    from brainfm_b200.misc import myzoom_torch_anisotropic
    y, aff = myzoom_torch_anisotropic(cu(GOLD["x"]), GOLD["aff"], [int(v) for v in GOLD["zoom_%s_size" % name]])
    assert np.array_equal(y.cpu().numpy(), GOLD["zoom_%s" % name])
    np.testing.assert_allclose(aff, GOLD["zoom_%s_aff" % name], rtol=1e-12, atol=1e-12)


def test_myzoom_torch_anisotropic_channels():
    from brainfm_b200.misc import myzoom_torch_anisotropic
    y = myzoom_torch_anisotropic(cu(GOLD["x4"]), None, [15, 8, 21])
    assert np.array_equal(y.cpu().numpy(), GOLD["zoom4"])


@pytest.mark.parametrize("name", ["r2", "r1", "r3"])
def test_torch_resize(name):
    from brainfm_b200.misc import torch_resize
    res = GOLD["resize_%s_res" % name]
    y, aff = torch_resize(cu(GOLD["x"]), GOLD["aff"], float(res) if res.ndim == 0 else res)
    want = GOLD["resize_%s" % name]
    assert tuple(y.shape) == want.shape
    np.testing.assert_allclose(y.cpu().numpy(), want, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(aff, GOLD["resize_%s_aff" % name], rtol=1e-12, atol=1e-12)


def test_torch_resize_channels():
    from brainfm_b200.misc import torch_resize
    y, aff = torch_resize(cu(GOLD["x4"]), GOLD["aff"], 2.0)
    np.testing.assert_allclose(y.cpu().numpy(), GOLD["resize4"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(aff, GOLD["resize4_aff"], rtol=1e-12, atol=1e-12)
    with pytest.raises(Exception):
        torch_resize(cu(GOLD["x"][0]), GOLD["aff"], 2.0)
