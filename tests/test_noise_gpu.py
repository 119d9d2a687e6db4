"""Distribution of the in-kernel counter-based normal generator (Philox4x32-10 + Box-Muller with the fast
`__log2f` / `__sincosf` intrinsics, csrc/common.cuh) -- the path bench.py times; the parity tests inject the
reference's draws instead.  Checked on >= 1.6e7 draws per stream: mean, variance, skewness, 4th and 6th moment,
tail mass, lag-1 correlation along every axis of a 160^3 x 4 volume, independence between streams and seeds, and
that the fused chain (GMM stage, k_gen_small) really draws this sequence."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _draw(n, seed, stream_id, first=0):
    from brainfm_b200 import _lib
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().bfm_philox_normal(out.data_ptr(), n, seed, stream_id, first, _stream()))
    return out


@pytest.mark.parametrize("stream_id", [0, 1, 2, 3])
def test_moments_and_tails(stream_id):
    n = 4 * 160 ** 3                                   # 1.6e7 draws
    x = _draw(n, 0x1234567 + 977 * stream_id, stream_id).double()
    assert torch.isfinite(x).all()
    m = x.mean().item()
    v = x.var().item()
    z = (x - m) / np.sqrt(v)
    skew, kurt, m6 = (z ** 3).mean().item(), (z ** 4).mean().item(), (z ** 6).mean().item()
    se = 1 / np.sqrt(n)
    assert abs(m) < 5 * se                             # mean: sd 1/sqrt(n)
    assert abs(v - 1) < 5 * np.sqrt(2) * se            # variance: sd sqrt(2/n)
    assert abs(skew) < 5 * np.sqrt(6) * se             # sd sqrt(6/n)
    assert abs(kurt - 3) < 5 * np.sqrt(96) * se        # sd sqrt(96/n)
    assert abs(m6 - 15) < 5 * np.sqrt(10170) * se      # var of z^6 = 10395 - 225
    # tail mass against the normal law (binomial standard deviations)
    from math import erfc, sqrt
    for t in (1.0, 2.0, 3.0, 4.0):
        p = erfc(t / sqrt(2))                          # two-sided
        got = (x.abs() > t).double().mean().item()
        assert abs(got - p) < 5 * sqrt(p * (1 - p) / n) + 1e-9, (t, got, p)
    assert x.abs().max().item() < 6.7                  # Box-Muller with u >= 2^-33: |x| <= sqrt(2*33*ln2) = 6.76
    assert x.abs().max().item() > 4.8                  # and the tails are populated (P(max < 4.8) ~ 1e-11)


@pytest.mark.parametrize("stream_id", [0, 1])
def test_no_voxel_to_voxel_correlation(stream_id):
    """Lag-1 (and lag-2, lag-4: inside / across a Philox group of four) autocorrelation along every axis."""
    n = 160
    x = _draw(n ** 3, 99, stream_id).view(n, n, n).double()
    bound = 5 / np.sqrt(x.numel())
    for axis in range(3):
        for lag in (1, 2, 3, 4):
            a = x.narrow(axis, 0, n - lag)
            b = x.narrow(axis, lag, n - lag)
            r = (a * b).mean().item()
            assert abs(r) < bound, (axis, lag, r)
    # squares too (Box-Muller pairs share a radius: cos / sin of the same angle are uncorrelated but dependent --
    # the dependence must not show up between NEIGHBOURING voxels' magnitudes beyond the pair itself)
    y = x * x - 1
    for lag in (2, 4):
        r = (y[..., :-lag] * y[..., lag:]).mean().item()
        assert abs(r) < 5 * 2 / np.sqrt(x.numel()), (lag, r)


def test_streams_and_seeds_are_independent():
    n = 1 << 22
    a, b, c = _draw(n, 7, 0).double(), _draw(n, 7, 1).double(), _draw(n, 8, 0).double()
    bound = 5 / np.sqrt(n)
    assert abs((a * b).mean().item()) < bound
    assert abs((a * c).mean().item()) < bound
    assert not torch.equal(a, b) and not torch.equal(a, c)
    # counter-based: a window of the sequence is the sequence
    w = _draw(4096, 7, 0, first=1000).double()
    assert torch.equal(w, a[4000:8096])


def test_uniformity_of_the_underlying_bits():
    """Probability-integral transform of the normals back to uniforms: chi-square over 256 bins."""
    n = 1 << 24
    x = _draw(n, 2024, 0).double()
    u = 0.5 * (1 + torch.erf(x / np.sqrt(2)))
    h = torch.histc(u, bins=256, min=0, max=1)
    chi2 = (((h - n / 256) ** 2) / (n / 256)).sum().item()
    assert chi2 < 255 + 6 * np.sqrt(2 * 255), chi2      # mean 255, sd sqrt(510)


def test_chain_draws_this_sequence():
    """The GMM stage with mu = 1000, sigma = 1 on a constant label map writes 1000 + eps: eps must be the stream-0
    sequence of the sample's seed at the absolute source voxel; k_gen_small must scale streams 2 / 3."""
    import bench
    from brainfm_b200 import _lib
    from tests import _inputs as ti
    size = 64
    shp = (size,) * 3
    old = bench.SIZE
    bench.SIZE = size
    try:
        subs = [dict(Gen=np.full(shp, 3, np.float32), T1=ti.smooth_image(shp, 0.0))]
        ds = bench.build_dataset(subs, torch.device("cuda", 0))
    finally:
        bench.SIZE = old
    np.random.seed(5)
    ds.generate_batch([0])
    torch.cuda.synchronize()
    descs, d_dev, B = ds._last_descs
    s = descs[0]
    assert s.eps_gmm is None and s.gen_small == 3
    # small grids: std * N(0,1) of streams 2 and 3
    nf = s.d.fs[0] * s.d.fs[1] * s.d.fs[2] * 3
    fs_view = torch.frombuffer(bytearray(4 * nf), dtype=torch.float32)            # host scratch
    _copy_from_device(fs_view, s.d.fsmall, 4 * nf)
    want = (_draw(nf, s.seed, 2) * s.fs_std).cpu()
    assert torch.equal(fs_view, want)
    nb = s.bs[0] * s.bs[1] * s.bs[2]
    bs_view = torch.frombuffer(bytearray(4 * nb), dtype=torch.float32)
    _copy_from_device(bs_view, s.bfsmall, 4 * nb)
    assert torch.equal(bs_view, (_draw(nb, s.seed, 3) * s.bf_std).cpu())
    # GMM stage: overwrite the tables with mu = 1000, sigma = 1 and re-run it on the same descriptors
    tab = torch.cat([torch.full((256,), 1000.0), torch.ones(256)]).cuda()
    copy = (_lib.GenSample * B)()
    C.memmove(C.addressof(copy), C.addressof(descs), C.sizeof(copy))
    copy[0].mu, copy[0].sigma = tab.data_ptr(), tab.data_ptr() + 1024
    copy[0].syn_pair_ok = 0
    dev = torch.from_numpy(np.frombuffer(copy, dtype=np.uint8).copy()).cuda()
    _lib.check(_lib.lib().bfm_gen_gmm(C.addressof(copy), dev.data_ptr(), B, _stream()))
    torch.cuda.synchronize()
    n = size ** 3
    syn = torch.frombuffer(bytearray(4 * n), dtype=torch.float32)
    _copy_from_device(syn, s.syn, 4 * n)
    eps = _draw(n, s.seed, 0).cpu()
    bb = ds._native.last_shapes()[0][0]
    got = syn.view(shp)[bb[0]:bb[3], bb[1]:bb[4], bb[2]:bb[5]]
    want = (1000.0 + eps).view(shp)[bb[0]:bb[3], bb[1]:bb[4], bb[2]:bb[5]]
    assert torch.equal(got, want)


def _copy_from_device(host_tensor, dev_ptr, nbytes):
    tmp = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    rt = C.CDLL("libcudart.so.12")
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    assert rt.cudaMemcpy(C.c_void_p(tmp.data_ptr()), C.c_void_p(dev_ptr), nbytes, 3) == 0       # device to device
    host_tensor.view(torch.uint8).copy_(tmp.cpu())
