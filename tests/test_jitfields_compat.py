"""brainfm_b200.interpol.jitfields_compat: the `jitfields`-shaped front end that lets the REFERENCE's own
utils.interpol forward to libbfm (utils/interpol/backend.py:1, jitfields.py:28-95).

CPU part (build container only -- needs /root/reference): the reference package is imported with our module
installed as `jitfields` and `backend.jitfields = True`; our kernels cannot run without a GPU, so the six entry
points of brainfm_b200.interpol.api are replaced by recorders that check the layout they receive and answer with the
reference's native implementation.  A call that goes reference API -> reference shim -> our front end -> (recorder)
-> back must then equal the reference's direct answer: that pins the argument marshalling in both directions.
GPU part: the front end against brainfm_b200.interpol.api itself (channels-last vs channels-first)."""
import importlib
import os
import sys
import tempfile

import numpy as np
import pytest
import torch

REF = "/root/reference/utils/interpol"


@pytest.fixture(scope="module")
def ref_interpol():
    if not os.path.isdir(REF):
        pytest.skip("reference checkout not present (GPU box)")
    import brainfm_b200.interpol.jitfields_compat as jf
    tmp = tempfile.mkdtemp(prefix="ref_interpol_")
    os.symlink(REF, os.path.join(tmp, "interpol"))          # not via utils/ (shadows logging)
    saved = {k: sys.modules.get(k) for k in ("jitfields", "interpol")}
    sys.modules["jitfields"] = jf
    sys.path.insert(0, tmp)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = importlib.import_module("interpol")
    assert mod.jitfields.available and mod.jitfields.jitfields is jf
    yield mod
    sys.path.remove(tmp)
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    for k in [k for k in sys.modules if k == "interpol" or k.startswith("interpol.")]:
        sys.modules.pop(k, None)


def _record_into_reference(monkeypatch, ref, calls):
    """brainfm_b200.interpol.api.* -> recorders that run the reference's native code path (backend off)."""
    from brainfm_b200.interpol import api

    def native(name):
        fn = getattr(ref, name)

        def run(*a, **k):
            calls.append((name, [tuple(x.shape) if torch.is_tensor(x) else x for x in a], dict(k)))
            ref.backend.jitfields = False
            try:
                return fn(*a, **k)
            finally:
                ref.backend.jitfields = True
        return run
    for name in ("grid_pull", "grid_push", "grid_count", "grid_grad", "spline_coeff", "spline_coeff_nd", "resize",
                 "restrict"):
        monkeypatch.setattr(api, name, native(name))


@pytest.mark.parametrize("order,bound", [(1, "dct2"), (3, "zero"), (2, "dft")])
def test_reference_forwards_to_the_front_end(ref_interpol, monkeypatch, order, bound):
    ref = ref_interpol
    torch.manual_seed(order)
    B, C, shp, oshp = 2, 3, (7, 6, 5), (4, 5, 6)
    x = torch.rand(B, C, *shp)
    grid = torch.rand(B, *oshp, 3) * torch.tensor([s - 1.0 for s in shp])
    gin = torch.rand(B, *shp, 3) * torch.tensor([s - 1.0 for s in oshp])
    kw = dict(interpolation=order, bound=bound, extrapolate=True)
    ref.backend.jitfields = False
    want = dict(pull=ref.grid_pull(x, grid, **kw), push=ref.grid_push(x, gin, oshp, **kw),
                count=ref.grid_count(gin, oshp, **kw), grad=ref.grid_grad(x, grid, **kw),
                coeff=ref.spline_coeff_nd(x, interpolation=3, bound="dct2", dim=3),
                resize=ref.resize(x, shape=[9, 8, 7], anchor='e', interpolation=order, bound=bound, prefilter=False))
    calls = []
    _record_into_reference(monkeypatch, ref, calls)
    ref.backend.jitfields = True
    try:
        got = dict(pull=ref.grid_pull(x, grid, **kw), push=ref.grid_push(x, gin, oshp, **kw),
                   count=ref.grid_count(gin, oshp, **kw), grad=ref.grid_grad(x, grid, **kw),
                   coeff=ref.spline_coeff_nd(x, interpolation=3, bound="dct2", dim=3),
                   resize=ref.resize(x, shape=[9, 8, 7], anchor='e', interpolation=order, bound=bound, prefilter=False))
    finally:
        ref.backend.jitfields = False
    names = [c[0] for c in calls]
    assert names == ["grid_pull", "grid_push", "grid_count", "grid_grad", "spline_coeff_nd", "resize"], names
    # what libbfm's API received: channel-first tensors again, the reference's keyword meaning
    assert calls[0][1] == [(B, C, *shp), (B, *oshp, 3)] and calls[0][2]["interpolation"] == order
    assert calls[1][1][:2] == [(B, C, *shp), (B, *shp, 3)] and tuple(calls[1][1][2]) == oshp
    assert calls[3][1] == [(B, C, *shp), (B, *oshp, 3)] and calls[3][2]["bound"] == bound
    for k in want:
        assert got[k].shape == want[k].shape, k
        assert torch.equal(got[k], want[k]), k


def test_unbatched_and_channel_free_inputs(ref_interpol, monkeypatch):
    """The reference's shim inserts a channel axis for bare spatial inputs (jitfields.py:12-18)."""
    ref = ref_interpol
    x = torch.rand(6, 5, 4)
    grid = torch.rand(3, 3, 3, 3) * 3
    ref.backend.jitfields = False
    want = ref.grid_pull(x, grid, interpolation=1, bound='dct2', extrapolate=True)
    calls = []
    _record_into_reference(monkeypatch, ref, calls)
    ref.backend.jitfields = True
    try:
        got = ref.grid_pull(x, grid, interpolation=1, bound='dct2', extrapolate=True)
    finally:
        ref.backend.jitfields = False
    assert calls[0][1] == [(1, 6, 5, 4), (3, 3, 3, 3)]
    assert torch.equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("order", [1, 3])
def test_front_end_equals_channel_first_api_on_gpu(order):
    import brainfm_b200.interpol as bi
    import brainfm_b200.interpol.jitfields_compat as jf
    torch.manual_seed(0)
    dev = "cuda"
    B, C, shp, oshp = 2, 3, (9, 8, 7), (5, 6, 4)
    x = torch.rand(B, C, *shp, device=dev)
    grid = torch.rand(B, *oshp, 3, device=dev) * torch.tensor([s - 1.0 for s in shp], device=dev)
    gin = torch.rand(B, *shp, 3, device=dev) * torch.tensor([s - 1.0 for s in oshp], device=dev)
    xl = torch.movedim(x, 1, -1).contiguous()
    kw = dict(bound='dct2', extrapolate=True)
    a = jf.pull(xl, grid, order=order, **kw)
    assert tuple(a.shape) == (B, *oshp, C)
    assert torch.equal(torch.movedim(a, -1, 1), bi.grid_pull(x, grid, interpolation=order, **kw))
    a = jf.push(xl, gin, oshp, order=order, **kw)
    assert tuple(a.shape) == (B, *oshp, C)
    np.testing.assert_allclose(torch.movedim(a, -1, 1).cpu().numpy(),
                               bi.grid_push(x, gin, oshp, interpolation=order, **kw).cpu().numpy(), rtol=1e-5, atol=1e-5)
    a = jf.grad(xl, grid, order=order, **kw)
    assert tuple(a.shape) == (B, *oshp, C, 3)
    assert torch.equal(torch.movedim(a, -2, 1), bi.grid_grad(x, grid, interpolation=order, **kw))
    np.testing.assert_allclose(jf.count(gin, oshp, order=order, **kw).cpu().numpy(),
                               bi.grid_count(gin, oshp, interpolation=order, **kw).cpu().numpy(), rtol=1e-5, atol=1e-5)
    assert torch.equal(jf.spline_coeff_nd(x, 3, bound='dct2', ndim=3), bi.spline_coeff_nd(x, 3, 'dct2', 3))
    out = torch.empty(B, *oshp, C, device=dev)
    assert jf.pull(xl, grid, order=order, out=out, **kw) is out
