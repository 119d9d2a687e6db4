"""TEST INFRASTRUCTURE ONLY.  Fixtures of the two resamplers the reference keeps in utils/misc.py
(myzoom_torch_anisotropic :1051-1115, torch_resize :1117-1187), produced by the unmodified reference on CPU.
-> tests/golden/misc.npz          python -m oracle.make_golden_misc   (build container only)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim as rs           # noqa: E402


def main():
    rs.install()
    from utils.misc import myzoom_torch_anisotropic, torch_resize
    rng = np.random.RandomState(3)
    gold = {}
    x = rng.rand(20, 24, 18).astype(np.float32)
    x4 = rng.rand(12, 10, 14, 3).astype(np.float32)
    aff = np.array([[1.2, 0.1, 0, -10], [0, 0.9, 0.2, 5], [0.05, 0, 2.5, 7], [0, 0, 0, 1]], dtype=np.float64)
    gold["x"], gold["x4"], gold["aff"] = x, x4, aff
    for name, newsize in (("up", [31, 29, 40]), ("down", [9, 11, 7]), ("mixed", [20, 37, 5])):
        y, a2 = myzoom_torch_anisotropic(torch.from_numpy(x), aff, newsize)
        gold["zoom_%s" % name], gold["zoom_%s_aff" % name] = y.numpy(), a2
        gold["zoom_%s_size" % name] = np.array(newsize)
    y = myzoom_torch_anisotropic(torch.from_numpy(x4), None, [15, 8, 21])
    gold["zoom4"] = y.numpy()
    with torch.no_grad():
        for name, res in (("r2", 2.0), ("r1", 1.0), ("r3", np.array([3.0, 1.0, 2.6]))):
            y, a2 = torch_resize(torch.from_numpy(x), aff, res, slow=True)
            gold["resize_%s" % name], gold["resize_%s_aff" % name] = y.numpy(), a2
            gold["resize_%s_res" % name] = np.asarray(res, dtype=np.float64)
        y, a2 = torch_resize(torch.from_numpy(x4), aff, 2.0, slow=True)
        gold["resize4"], gold["resize4_aff"] = y.numpy(), a2
    gold["meta.versions"] = np.array("torch %s numpy %s" % (torch.__version__, np.__version__))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "misc.npz"), **gold)
    print({k: v.shape for k, v in gold.items() if hasattr(v, "shape") and v.ndim >= 3})


if __name__ == "__main__":
    main()
