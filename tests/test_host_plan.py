"""Host-side planning logic (no GPU): banded blur-o-downsample maps and zoom tables against the oracle."""
import numpy as np
import pytest
import torch

from brainfm_b200 import plan
from oracle import gen_oracle as go


def _apply_band(x, axis, start, w):
    n_out, T = w.shape
    shp = list(x.shape)
    n_in = shp[axis]
    shp[axis] = n_out
    out = np.zeros(shp, dtype=np.float64)
    xm = np.moveaxis(x.astype(np.float64), axis, 0)
    om = np.moveaxis(out, axis, 0)
    for q in range(n_out):
        for t in range(T):
            s = start[q] + t
            if 0 <= s < n_in:
                om[q] += float(w[q, t]) * xm[s]
    return out


@pytest.mark.parametrize("new,stds", [((20, 20, 7), (0.0, 0.0, 1.7)), ((20, 6, 20), (0.0, 2.4, 0.0)),
                                      ((9, 11, 5), (1.2, 0.9, 2.95)), ((20, 20, 20), (0.0, 0.0, 0.0)),
                                      ((13, 20, 10), (0.0, 0.0, 0.0))])
def test_band_maps_equal_blur_then_trilinear(new, stds):
    torch.manual_seed(0)
    size = (20, 20, 20)
    vol = torch.rand(size) * 100
    B = go.blur3d(vol, np.array(stds))
    fac = np.array(new) / np.array(size)
    v = [go.zoom_tables(size[d], fac[d], new[d], dtype64=True) for d in range(3)]
    II, JJ, KK = torch.meshgrid(*v, indexing="ij")
    ref = go.sample_trilinear(B, II, JJ, KK).numpy()
    x = vol.numpy()
    for ax in range(3):
        start, w, T = plan.band_host(size[ax], new[ax], stds[ax])
        assert w.shape == (new[ax], T)
        x = _apply_band(x, ax, start, w)
    np.testing.assert_allclose(x, ref, rtol=1e-5, atol=1e-4)
    if new[0] == size[0] and stds[0] == 0:
        assert np.all(x[0] == 0)          # strict `>0` mask zeroes the first plane of identity axes


def test_zoom_tables_match_oracle():
    for n_in, n_out in [(5, 160), (9, 160), (25, 160), (160, 160), (3, 64)]:
        f = n_out / n_in
        a = plan.zoom_tables_host(n_in, f, n_out)
        b = go.zoom_tables(n_in, f, n_out)
        for x, y in zip(a, b):
            assert np.array_equal(np.asarray(x), y.numpy())


def test_c_abi_exports_every_declared_symbol():
    import re, os
    from brainfm_b200 import _lib
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "bfm.h")).read()
    declared = set(re.findall(r"\b(bfm_[a-z0-9_]+)\s*\(", hdr))
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(_lib.exported_symbols())
    assert L.bfm_abi_version() == _lib.ABI_VERSION == 8


@pytest.mark.parametrize("seed", range(6))
def test_bbox_candidates_bound_the_full_scan(seed):
    """The candidate-voxel bounding box (bfm_gen_bbox) relies on: extrema of the clamped source coordinates
    over ALL voxels lie within 2*delta of the extrema over the candidate voxels (delta = 1e-3)."""
    rng = np.random.RandomState(seed)
    torch.manual_seed(seed)
    size = [48, 40, 56]
    src = [64, 64, 64]
    fs = [int(rng.randint(2, 6)) for _ in range(3)]
    Fsmall = float(rng.rand() * 4) * torch.randn(*fs, 3)
    F = go.zoom_linear(Fsmall, np.array(size) / np.array(fs))
    rot = (rng.rand(3) * 30 - 15) / 180 * np.pi
    A = torch.tensor(go.affine_matrix(rot, rng.rand(3) * 0.4 - 0.2, 1 + rng.rand(3) * 0.4 - 0.2), dtype=torch.float32)
    c2 = torch.tensor((np.array(src) - 1) / 2, dtype=torch.float32)
    _, centred = go.centred_grid(size)
    p = [centred[d] + F[..., d] for d in range(3)]
    cands = [torch.from_numpy(plan.zoom_candidates_host(fs[a], size[a] / fs[a], size[a]).astype(np.int64))
             for a in range(3)]
    for r in range(3):
        v = A[r, 0] * p[0] + A[r, 1] * p[1] + A[r, 2] * p[2] + c2[r]
        v = v.clamp(0, src[r] - 1)
        sub = v[cands[0]][:, cands[1]][:, :, cands[2]]
        assert 0 <= float(sub.min() - v.min()) <= 2e-3
        assert 0 <= float(v.max() - sub.max()) <= 2e-3
