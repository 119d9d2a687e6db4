"""TEST INFRASTRUCTURE ONLY.  Runs the reference's utils/interpol (torch CPU) on small seeded problems and stores
inputs + outputs in tests/golden/interpol.npz; checks oracle/interpol_oracle.py against it on the spot.

    python -m oracle.make_golden_interpol      (build container only)
"""
import itertools
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim as rs          # noqa: E402
from oracle import interpol_oracle as io   # noqa: E402

BOUND_NAMES = ['zero', 'replicate', 'dct1', 'dct2', 'dst1', 'dst2', 'dft']


def cases():
    """(key, kind, dim, order, bound, extrapolate)"""
    out = []
    for dim in (1, 2, 3):
        orders = range(0, 4) if dim == 3 else (0, 1, 3)
        for o, b, e in itertools.product(orders, range(7), (0, 1, 2)):
            if dim < 3 and e == 2 and b not in (0, 3):
                continue
            for kind in ('pull', 'push', 'count', 'grad'):
                out.append(("%s_d%d_o%d_b%d_e%d" % (kind, dim, o, b, e), kind, dim, [o] * dim, [b] * dim, e))
    for o in (4, 5, 6, 7):                           # high orders (dct2, dft)
        for b in (3, 6):
            for kind in ('pull', 'grad', 'push'):
                out.append(("%s_d3_o%d_b%d_e1" % (kind, o, b), kind, 3, [o] * 3, [b] * 3, 1))
    # mixed orders / bounds (generic nd path incl. order 0 with floor(g+0.5))
    out.append(("pull_d3_mixed_a", 'pull', 3, [0, 1, 3], [1, 3, 6], 1))
    out.append(("pull_d3_mixed_b", 'pull', 3, [3, 0, 2], [5, 4, 2], 0))
    out.append(("grad_d3_mixed_a", 'grad', 3, [1, 3, 2], [3, 0, 6], 2))
    out.append(("push_d3_mixed_a", 'push', 3, [2, 1, 0], [6, 3, 1], 1))
    return out


def problem(dim, seed):
    rng = np.random.RandomState(seed)
    ishape = [7, 6, 5][:dim]
    oshape = [6, 5, 4][:dim]
    vol = rng.randn(2, 2, *ishape).astype(np.float32)
    grid = np.stack([rng.uniform(-4, n + 3, size=[2, *oshape]) for n in ishape], -1).astype(np.float32)
    # adversarial coordinates: exact integers, halves, thresholds
    flat = grid.reshape(2, -1, dim)
    special = [0.0, -0.05, -0.04, 0.5, 1.5, 2.5, -0.5, -0.55, -0.56, ishape[0] - 1.0, ishape[0] - 0.96, ishape[0] - 0.5]
    for i, v in enumerate(special[:flat.shape[1]]):
        flat[0, i, 0] = v
    push_in = rng.randn(2, 2, *oshape).astype(np.float32)
    return vol, grid, push_in, ishape, oshape


def main():
    rs.install()
    import utils.interpol as ri
    gold = {}
    worst = 0.0
    for key, kind, dim, order, bound, e in cases():
        vol, grid, push_in, ishape, oshape = problem(dim, 100 + dim)
        tv, tg, tp = torch.from_numpy(vol), torch.from_numpy(grid), torch.from_numpy(push_in)
        bn = [BOUND_NAMES[b] for b in bound]
        if kind == 'pull' and dim == 2 and max(order) == 0 and e != 1:
            continue        # reference bug: iso0.pull2d computes `mask * mask` (iso0.py:155) and then fails to reshape
        if kind == 'pull':
            ref = ri.grid_pull(tv, tg, order, bn, e).numpy()
            orc = io.pull(vol, grid, order, bound, e) if max(order) <= 3 else None
        elif kind == 'grad':
            ref = ri.grid_grad(tv, tg, order, bn, e).numpy()
            orc = io.pull(vol, grid, order, bound, e, grad=True) if max(order) <= 3 else None
        elif kind == 'push':
            ref = ri.grid_push(tp, tg, ishape, order, bn, e).numpy()
            orc = io.push(push_in, grid, ishape, order, bound, e) if max(order) <= 3 else None
        else:
            ref = ri.grid_count(tg, ishape, order, bn, e).numpy()
            orc = io.push(None, grid, ishape, order, bound, e) if max(order) <= 3 else None
        gold[key] = ref
        if orc is not None:
            d = float(np.abs(ref - orc).max())
            worst = max(worst, d)
            assert d < 2e-5, (key, d)
    # prefilter: orders 2..7, bounds zero/replicate/dct1/dct2/dft, short and long lines
    rng = np.random.RandomState(7)
    for n in (5, 40):
        x = rng.randn(3, n, 4).astype(np.float32)
        gold["coeff_in_n%d" % n] = x
        for o in range(2, 8):
            for b in (0, 1, 2, 3, 6):
                ref = ri.spline_coeff(torch.from_numpy(x), o, BOUND_NAMES[b], dim=1).numpy()
                gold["coeff_n%d_o%d_b%d" % (n, o, b)] = ref
                if o <= 3:
                    d = float(np.abs(ref - io.spline_filter(x, o, b, 1)).max())
                    worst = max(worst, d)
                    assert d < 2e-5, (n, o, b, d)
    v3 = rng.randn(2, 6, 7, 8).astype(np.float32)
    gold["coeffnd_in"] = v3
    gold["coeffnd_o3_dct2"] = ri.spline_coeff_nd(torch.from_numpy(v3), 3, 'dct2', dim=3).numpy()
    # resize / restrict
    img = rng.rand(1, 2, 6, 5, 4).astype(np.float32)
    gold["resize_in"] = img
    for anchor in ('c', 'e', 'f', 'l'):
        kw = dict(factor=[2, 1.5, 2]) if anchor in ('f', 'l') else dict(shape=[12, 8, 8])
        gold["resize_%s_o1" % anchor] = ri.resize(torch.from_numpy(img), anchor=anchor, interpolation=1, **kw).numpy()
        gold["resize_%s_o3" % anchor] = ri.resize(torch.from_numpy(img), anchor=anchor, interpolation=3, bound='dct2',
                                                  **kw).numpy()
        kw = dict(factor=[2, 1.5, 2]) if anchor in ('f', 'l') else dict(shape=[3, 3, 2])
        gold["restrict_%s_o1" % anchor] = ri.restrict(torch.from_numpy(img), anchor=anchor, interpolation=1, **kw).numpy()
    # the call the generator makes (datasets.py:338): cubic dct2 prefiltered edge-anchored resize
    low = rng.rand(10, 12, 4).astype(np.float32) * 100
    gold["bspline_zoom_in"] = low
    gold["bspline_zoom_out"] = ri.resize(torch.from_numpy(low), shape=[20, 24, 16], anchor='edge', interpolation=3,
                                         bound='dct2', prefilter=True).numpy()
    # SURVEY appendix C known answers (1-D)
    x = torch.tensor([1., 2., 3., 4., 5.])
    g = torch.arange(-4., 9.)[:, None]
    for b in BOUND_NAMES:
        for o in (0, 1, 3):
            gold["appC_%s_o%d" % (b, o)] = ri.grid_pull(x, g, o, b, True).numpy()
    # label (integer) pull
    lab = torch.from_numpy(rng.randint(0, 5, size=(6, 5, 4))).to(torch.int64)
    gl = torch.from_numpy(np.stack([rng.uniform(-1, n, size=(5, 4, 3)) for n in (6, 5, 4)], -1).astype(np.float32))
    gold["label_in"], gold["label_grid"] = lab.numpy(), gl.numpy()
    gold["label_out"] = ri.grid_pull(lab, gl, 1, 'dct2', True).numpy()
    gold["meta.versions"] = np.array("torch %s numpy %s" % (torch.__version__, np.__version__))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "interpol.npz"), **gold)
    print("interpol fixtures: %d arrays; oracle-vs-reference worst |diff| = %.3e" % (len(gold), worst))


if __name__ == "__main__":
    main()
