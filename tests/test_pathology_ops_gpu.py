"""Op-level kernels of the pathology branch (csrc/pathology.cu) against the reference's tensor expressions
(Generator/datasets.py:364-372, 391-400, 496-518) evaluated with torch on the same device."""
import ctypes as C

import numpy as np
import pytest
import torch

from brainfm_b200 import _lib

pytestmark = pytest.mark.gpu


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("u8", [True, False])
def test_gmm_crop_matches_the_tensor_expression(u8):
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(3)
    src = (21, 34, 28)
    lab = torch.randint(0, 256, src, generator=g)
    lab[2:5, 3:9, 4:8] = 77
    box = (3, 5, 2, 19, 30, 27)
    mus, sigmas = torch.rand(256, generator=g) * 200 + 25, torch.rand(256, generator=g) * 20 + 5
    crop = tuple(box[3 + a] - box[a] for a in range(3))
    eps = torch.randn(crop, generator=g)
    # reference expression
    G = lab[box[0]:box[3], box[1]:box[4], box[2]:box[5]].float()
    G[G == 77] = 2
    Gr = torch.round(G).long()
    ref = mus[Gr] + sigmas[Gr] * eps
    ref[ref < 0] = 0
    L = _lib.lib()
    d_lab = (lab.to(torch.uint8) if u8 else lab.float()).to(dev).contiguous()
    out = torch.empty(crop, dtype=torch.float32, device=dev)
    d_eps, d_mu, d_sg = eps.to(dev), mus.to(dev), sigmas.to(dev)
    _lib.check(L.bfm_gmm_crop(d_lab.data_ptr(), int(u8), (C.c_int * 3)(*src), (C.c_int * 6)(*box), d_mu.data_ptr(),
                              d_sg.data_ptr(), d_eps.data_ptr(), 0, out.data_ptr(), _st()))
    assert torch.equal(out.cpu(), ref)
    # counter-based noise: same field as the fused chain's (stream 0, keyed on the absolute source voxel)
    _lib.check(L.bfm_gmm_crop(d_lab.data_ptr(), int(u8), (C.c_int * 3)(*src), (C.c_int * 6)(*box), d_mu.data_ptr(),
                              d_sg.data_ptr(), 0, 1234, out.data_ptr(), _st()))
    n_src = int(np.prod(src))
    field = torch.empty((n_src + 3) // 4 * 4, dtype=torch.float32, device=dev)
    _lib.check(L.bfm_philox_normal(field.data_ptr(), field.numel(), 1234, 0, 0, _st()))
    e2 = field[:n_src].view(src)[box[0]:box[3], box[1]:box[4], box[2]:box[5]].cpu()
    ref2 = mus[Gr] + sigmas[Gr] * e2
    ref2[ref2 < 0] = 0
    assert torch.equal(out.cpu(), ref2)


def test_cerebral_mask_and_tissue_means():
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(5)
    src = (24, 24, 24)
    lab = torch.randint(0, 60, src, generator=g).to(torch.uint8)
    box = (0, 0, 0, 24, 24, 24)
    syn = torch.rand(src, generator=g) * 100
    Gr = lab.long()
    cer = syn.clone()
    cer[Gr == 0] = 0
    wm = (Gr == 2) | (Gr == 41)
    gm = (Gr != 0) & (Gr != 2) & (Gr != 41)
    L = _lib.lib()
    d_syn, d_lab = syn.to(dev), lab.to(dev)
    out = torch.empty_like(d_syn)
    sums = torch.empty(4, dtype=torch.float64, device=dev)
    _lib.check(L.bfm_pathol_cerebral(d_syn.data_ptr(), d_lab.data_ptr(), 1, (C.c_int * 3)(*src), (C.c_int * 6)(*box),
                                     out.data_ptr(), sums.data_ptr(), _st()))
    assert torch.equal(out.cpu(), cer)
    s = sums.cpu().numpy()
    np.testing.assert_allclose(s, [float((syn.double() * wm).sum()), float(wm.sum()), float((syn.double() * gm).sum()),
                                   float(gm.sum())], rtol=1e-12)
    # P[C == 0] = 0 for float64 and float32 maps
    for dt in (torch.float64, torch.float32):
        P = torch.rand(src, generator=g, dtype=torch.float64).to(dt)
        want = P.clone()
        want[cer == 0] = 0
        d_P = P.to(dev)
        _lib.check(L.bfm_zero_where_zero(d_P.data_ptr(), int(dt == torch.float64), out.data_ptr(), d_P.numel(), _st()))
        assert torch.equal(d_P.cpu(), want)


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("direction", [True, False])
def test_encode_pathology_matches_the_tensor_expression(dt, direction):
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(11)
    shape = (20, 30, 26)
    I = torch.rand(shape, generator=g) * 150
    P = (torch.rand(shape, generator=g, dtype=torch.float64) ** 3).to(dt)
    Pprob = torch.rand(shape, generator=g, dtype=torch.float64).to(dt)
    eps = torch.randn(shape, generator=g)
    u_mu, u_sg = torch.rand(10000, generator=g), torch.rand(10000, generator=g)
    # reference (datasets.py:496-518), on the CPU
    I_ref = I.clone()
    I_mu = (I_ref * P).sum() / P.sum()
    p_mask = torch.round(P).long()
    pth_mus = 3 * I_mu / 4 + I_mu / 4 * u_mu
    pth_mus = pth_mus if direction else -pth_mus
    pth_sigmas = I_mu / 4 * u_sg
    I_ref += Pprob * (pth_mus[p_mask] + pth_sigmas[p_mask] * eps)
    I_ref[I_ref < 0] = 0
    # kernels
    L = _lib.lib()
    d_I, d_P, d_Pp, d_eps = I.to(dev), P.to(dev), Pprob.to(dev), eps.to(dev)
    sums = torch.empty(2, dtype=torch.float64, device=dev)
    dbl = int(dt == torch.float64)
    _lib.check(L.bfm_masked_mean(d_I.data_ptr(), d_P.data_ptr(), dbl, d_I.numel(), sums.data_ptr(), _st()))
    mu_dev = float(sums[0] / sums[1])
    assert abs(mu_dev - float(I_mu)) <= (1e-12 if dbl else 2e-6) * abs(float(I_mu))
    # feed the reference's tables so that the element-wise kernel is compared bit for bit
    d_mus, d_sgs = pth_mus.float().to(dev), pth_sigmas.float().to(dev)
    _lib.check(L.bfm_encode_pathology(d_I.data_ptr(), d_P.data_ptr(), d_Pp.data_ptr(), dbl, d_mus.data_ptr(),
                                      d_sgs.data_ptr(), 10000, d_eps.data_ptr(), 0, d_I.numel(), _st()))
    assert torch.equal(d_I.cpu(), I_ref)
