"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference generator path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module, and only as the checker / CPU baseline.  The product package
(brainfm_b200/) never imports anything under oracle/.

What it is: a torch-CPU / numpy restatement of jhuldr/BrainFM's `BaseGen.__getitem__` chain
(Generator/datasets.py, Generator/utils.py), written independently (vectorised, flat-index
gathers instead of advanced indexing, no Python loops over planes) but performing the SAME
IEEE-754 fp32 operations in the SAME order, and consuming the numpy / python / torch global RNGs
in the SAME order as the reference, so that for identical seeds it reproduces the reference's
outputs.  Each function cites the reference file:line it follows.

Pinning: tests/test_oracle_vs_golden.py compares this module against tests/golden/*.npz,
which were produced by running the UNMODIFIED reference in the build container
(oracle/make_golden.py, through oracle/ref_shim.py).  The reference ships no golden vectors
for this path, so those fixtures are the pin.

Every random draw is appended to `self.log` as (tag, value) so that the CUDA path can be fed
the identical draws (brainfm_b200.draws.ReplayDraws).
"""
import math
import random as pyrandom

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# constants (Generator/constants.py:284-289, Generator/utils.py:664-669)
# ----------------------------------------------------------------------------------------------
LABELS_BRAINSEG_EXTRACEREBRAL = [0, 11, 12, 13, 16, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 46,
                                 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 14, 15, 17, 47, 49, 51, 53, 55,
                                 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 48, 50, 52, 54, 56]
N_NEUTRAL = 20
LABELS_BRAINSEG_LEFT = [0, 1, 2, 3, 4, 7, 8, 9, 10, 14, 15, 17, 31, 34, 36, 38, 40, 42]       # constants.py:289
CT_GROUPS = {
    "darker": [4, 5, 14, 15, 24, 31, 72],
    "dark": [2, 7, 16, 77, 30],
    "bright": [3, 8, 17, 18, 28, 10, 11, 12, 13, 26],
    "brighter": [],
}


# ----------------------------------------------------------------------------------------------
# samplers
# ----------------------------------------------------------------------------------------------
def _as4(X):
    return X if X.dim() == 4 else X.unsqueeze(-1)


def sample_nearest(X, I, J, K):
    """Nearest-neighbour gather: round-half-even then clamp (Generator/utils.py:124-138)."""
    X4 = _as4(X)
    nx, ny, nz, C = X4.shape
    ri = torch.round(I).long().clamp_(0, nx - 1)
    rj = torch.round(J).long().clamp_(0, ny - 1)
    rk = torch.round(K).long().clamp_(0, nz - 1)
    flat = (ri * ny + rj) * nz + rk
    out = X4.reshape(-1, C)[flat.reshape(-1)].reshape(*I.shape, C)
    return out[..., 0] if C == 1 else out


def sample_trilinear(X, I, J, K, default=0.0):
    """Trilinear gather with the strict `>0` / `<=n-1` validity mask, ceil-clamp and
    x->y->z lerp order (Generator/utils.py:140-192)."""
    X4 = _as4(X)
    nx, ny, nz, C = X4.shape
    ok = (I > 0) & (J > 0) & (K > 0) & (I <= nx - 1) & (J <= ny - 1) & (K <= nz - 1)
    okf = ok.reshape(-1)
    iv, jv, kv = I.reshape(-1)[okf], J.reshape(-1)[okf], K.reshape(-1)[okf]

    def split(v, n):
        lo = torch.floor(v).long()
        hi = (lo + 1).clamp_(max=n - 1)
        wh = (v - lo).unsqueeze(-1)
        return lo, hi, 1 - wh, wh

    x0, x1, ax0, ax1 = split(iv, nx)
    y0, y1, ay0, ay1 = split(jv, ny)
    z0, z1, az0, az1 = split(kv, nz)
    Xf = X4.reshape(-1, C)

    def at(a, b, c):
        return Xf[(a * ny + b) * nz + c]

    # x first (four edges), then y, then z -- each `*` and `+` separately rounded in fp32
    e00 = at(x0, y0, z0) * ax0 + at(x1, y0, z0) * ax1
    e01 = at(x0, y0, z1) * ax0 + at(x1, y0, z1) * ax1
    e10 = at(x0, y1, z0) * ax0 + at(x1, y1, z0) * ax1
    e11 = at(x0, y1, z1) * ax0 + at(x1, y1, z1) * ax1
    f0 = e00 * ay0 + e10 * ay1
    f1 = e01 * ay0 + e11 * ay1
    val = f0 * az0 + f1 * az1

    out = torch.zeros(I.numel(), C, dtype=torch.float32)
    out[okf] = val.float()
    out[~okf] = default
    out = out.reshape(*I.shape, C)
    return out[..., 0] if C == 1 else out


def surface_deform(mat, A, c2, Fneg, flip, size):
    """read_and_deform_surface (Generator/utils.py:479-533): vertices through the inverse affine and the integrated
    negative field; x mirrored and left / right swapped on flip; faces untouched."""
    Ainv = torch.inverse(A)
    out = {}
    for v, f in (("Vlw", "Flw"), ("Vrw", "Frw"), ("Vlp", "Flp"), ("Vrp", "Frp")):
        V = torch.tensor(np.asarray(mat[v]), dtype=torch.float)
        V = V - c2[None, :]
        V = V @ torch.transpose(Ainv, 0, 1)
        V = V + sample_trilinear(Fneg, V[:, 0] + c2[0], V[:, 1] + c2[1], V[:, 2] + c2[2])
        V = V + c2[None, :]
        out[v] = V
        out[f] = torch.tensor(np.asarray(mat[f]), dtype=torch.int)
    if flip:
        for v in ("Vlw", "Vrw", "Vlp", "Vrp"):
            out[v][:, 0] = size[0] - 1 - out[v][:, 0]
        for a, b in (("Vlw", "Vrw"), ("Vlp", "Vrp"), ("Flw", "Frw"), ("Flp", "Frp")):
            out[a], out[b] = out[b], out[a]
    return out


def zoom_tables(n_in, factor, n_out, dtype64=False):
    """1-D coordinate / index / weight tables of the separable zoom (Generator/utils.py:205-236).
    float32 `torch.arange` by default; the float64 numpy variant is the one
    resample_resolution uses (Generator/utils.py:597-601)."""
    delta = (1.0 - factor) / (2.0 * factor)
    if dtype64:
        v = np.arange(delta, delta + n_out / factor, 1 / factor)[:n_out]
        return torch.tensor(v, dtype=torch.float32)
    v = torch.arange(delta, delta + n_out / factor, 1 / factor, dtype=torch.float32)[:n_out]
    v = v.clamp(min=0, max=n_in - 1)
    lo = torch.floor(v).int()
    hi = (lo + 1).clamp(max=n_in - 1)
    wh = v - lo
    return lo.long(), hi.long(), 1 - wh, wh


def zoom_linear(X, factor):
    """Separable, edge-clamped linear zoom, axis 0 then 1 then 2 (Generator/utils.py:200-257)."""
    factor = np.asarray(factor, dtype=np.float64)
    squeeze = X.dim() == 3
    Y = _as4(X)
    newsize = np.round(np.array(Y.shape[:-1]) * factor).astype(int)
    for ax in range(3):
        lo, hi, wl, wh = zoom_tables(Y.shape[ax], factor[ax], int(newsize[ax]))
        shp = [1, 1, 1, 1]
        shp[ax] = -1
        Y = wl.reshape(shp) * Y.index_select(ax, lo) + wh.reshape(shp) * Y.index_select(ax, hi)
    return Y[..., 0] if squeeze else Y


# ----------------------------------------------------------------------------------------------
# geometry
# ----------------------------------------------------------------------------------------------
def bspline_resize(I, size):
    """interpol.resize(I, shape=size, anchor='edge', interpolation=3, bound='dct2', prefilter=True)
    (Generator/datasets.py:337-338; utils/interpol/resize.py:74-128, api.py:137-212): cubic prefilter along every
    axis (dct2), then a cubic pull at arange(out) * (in / out) + 0.5 * (in / out - 1), float32 throughout."""
    from oracle import interpol_oracle as io
    x = I.numpy().astype(np.float32)
    for axis in range(3):
        x = io.spline_filter(x, 3, 3, axis)
    lin = []
    for n_in, n_out in zip(x.shape, size):
        scale = n_in / n_out
        lin.append((torch.arange(0., n_out, dtype=torch.float32) * scale + 0.5 * (scale - 1)).numpy())
    grid = np.stack(np.meshgrid(*lin, indexing="ij"), axis=-1)[None].astype(np.float32)
    out = io.pull(x[None, None], grid, [3, 3, 3], [3, 3, 3], 1)
    return torch.from_numpy(np.ascontiguousarray(out[0, 0]))


def affine_matrix(rot, sh, s):
    """SHx.SHy.SHz.Rx.Ry.Rz with rows scaled by s, float64 (Generator/utils.py:102-116)."""
    c, n = np.cos(rot), np.sin(rot)
    Rx = np.array([[1, 0, 0], [0, c[0], -n[0]], [0, n[0], c[0]]])
    Ry = np.array([[c[1], 0, n[1]], [0, 1, 0], [-n[1], 0, c[1]]])
    Rz = np.array([[c[2], -n[2], 0], [n[2], c[2], 0], [0, 0, 1]])
    SHx = np.array([[1, 0, 0], [sh[1], 1, 0], [sh[2], 0, 1]])
    SHy = np.array([[1, sh[0], 0], [0, 1, 0], [0, sh[2], 1]])
    SHz = np.array([[1, 0, sh[0]], [0, 1, sh[1]], [0, 0, 1]])
    A = SHx @ SHy @ SHz @ Rx @ Ry @ Rz
    return A * np.asarray(s, dtype=np.float64)[:, None]


def centred_grid(size):
    """xc,yc,zc = index - (size-1)/2 in fp32 (Generator/datasets.py:146-162)."""
    ax = [torch.arange(n, dtype=torch.float32) for n in size]
    c = torch.tensor((np.array(size) - 1) / 2, dtype=torch.float32)
    g = torch.meshgrid(*ax, indexing="ij")
    return g, tuple(g[d] - c[d] for d in range(3))


def deform_grid(centred, shp, A, c2, Fld):
    """Coordinates of every output voxel in the source volume, clamped, plus the bounding box
    and the bbox-relative coordinates (Generator/datasets.py:264-303)."""
    if Fld is not None:
        p = [centred[d] + Fld[..., d] for d in range(3)]
    else:
        p = list(centred)
    q = []
    for r in range(3):
        v = A[r, 0] * p[0] + A[r, 1] * p[1] + A[r, 2] * p[2] + c2[r]
        v = torch.where(v < 0, torch.zeros_like(v), v)
        hi = float(shp[r] - 1)
        v = torch.where(v > hi, torch.full_like(v, hi), v)
        q.append(v)
    lo = [torch.floor(torch.min(v)) for v in q]
    hi = [1 + torch.ceil(torch.max(v)) for v in q]
    rel = [q[d] - lo[d] for d in range(3)]
    lo_i = [int(v.item()) for v in lo]
    hi_i = [int(v.item()) for v in hi]
    return rel, lo_i, hi_i


def svf_integrate(Fld, grid, n_steps):
    """Scaling and squaring of a stationary velocity field (Generator/datasets.py:214-223)."""
    G = Fld * (1.0 / (2.0 ** n_steps))
    for _ in range(n_steps):
        G = G + sample_trilinear(G, grid[0] + G[..., 0], grid[1] + G[..., 1], grid[2] + G[..., 2])
    return G


# ----------------------------------------------------------------------------------------------
# resolution degradation
# ----------------------------------------------------------------------------------------------
def gaussian_taps(sigma):
    """Normalised taps on [-ceil(3s), ceil(3s)] (Generator/utils.py:74-81)."""
    half = int(np.ceil(3 * sigma))
    t = torch.linspace(-half, half, 2 * half + 1, dtype=torch.float32)
    g = torch.exp(-(t / sigma) ** 2 / 2)
    return g / g.sum()


def blur3d(vol, stds):
    """Separable zero-padded Gaussian correlation, axes 0,1,2, skipped where std==0
    (Generator/utils.py:83-94)."""
    v = vol[None, None]
    for ax in range(3):
        if stds[ax] > 0:
            k = gaussian_taps(stds[ax])
            kshape = [1, 1, 1, 1, 1]
            kshape[2 + ax] = -1
            pad = [0, 0, 0]
            pad[ax] = len(k) // 2
            v = F.conv3d(v, k.reshape(kshape), stride=1, padding=tuple(pad))
    return torch.squeeze(v)


def draw_resolution(low_res_only, log):
    """Acquisition resolution / slice thickness classes (Generator/utils.py:34-57)."""
    r = np.random.rand() * 0.5 + 0.5 if low_res_only else np.random.rand()
    log.append(("res.class", r))
    res = np.ones(3)
    thick = np.ones(3)
    if r < 0.25:
        pass
    elif r < 0.5:
        ax = np.random.randint(3)
        u = np.random.rand()
        u2 = np.random.rand()
        log.append(("res.axis", ax)); log.append(("res.u", u)); log.append(("res.u2", u2))
        res[ax] = 2.5 + 6 * u
        thick[ax] = np.min([res[ax], 4.0 + 2.0 * u2])
    elif r < 0.75:
        u = np.random.rand(3)
        log.append(("res.u3", u))
        res = np.array([1.3, 1.3, 4.8]) + 0.4 * u
        thick = res.copy()
    else:
        u = np.random.rand(3)
        log.append(("res.u3", u))
        res = 2.0 + 3.0 * u
        thick = res.copy()
    return res, thick


# ----------------------------------------------------------------------------------------------
# the generator
# ----------------------------------------------------------------------------------------------
class GeneratorOracle:
    """Restates BaseGen / BrainIDGen (Generator/datasets.py:25-757) over in-memory volumes.

    `cfg` is the reference's Namespace tree (or any object with the same attributes).
    `volumes` maps modality keys to numpy arrays:
      'Gen' (generation labels, required), 'T1' (required: always read, datasets.py:657),
      optional 'T2','FLAIR','CT','segmentation','distance' (list of 4),'registration' (list of 3),
      'bias_field'.
    """

    def __init__(self, cfg, volumes, dataset_name="HCP", case_name="HCP.sub01", brain_id=False):
        self.cfg = cfg
        self.g = cfg.generator
        self.vol = volumes
        self.dataset_name = dataset_name
        self.case_name = case_name
        self.brain_id = brain_id
        self.size = list(self.g.size)
        self.res = np.array([1.0, 1.0, 1.0])
        self.grid, self.centred = centred_grid(self.size)
        self.tasks = [k for k, v in vars(cfg.task).items() if v]
        if "bias_field" in self.tasks and "segmentation" not in self.tasks:
            self.tasks.append("segmentation")          # datasets.py:125-127
        # datasets.py:165-171: the left-hemisphere label list when left_hemis_only
        labels = LABELS_BRAINSEG_LEFT if getattr(self.g, "left_hemis_only", False) else LABELS_BRAINSEG_EXTRACEREBRAL
        self.hemis_mask = None
        self.lut = torch.zeros(10000, dtype=torch.long)
        for i, l in enumerate(labels):
            self.lut[l] = i
        self.eye = torch.eye(len(labels), dtype=torch.float32)
        nlat = (len(labels) - N_NEUTRAL) // 2
        self.vflip = np.concatenate([np.arange(N_NEUTRAL), np.arange(N_NEUTRAL + nlat, len(labels)),
                                     np.arange(N_NEUTRAL, N_NEUTRAL + nlat)])
        self.aug_steps = vars(cfg.augmentation_steps)
        self.log = []

    # -- helpers -------------------------------------------------------------------------------
    def _merge(self, ns):
        for k, v in vars(ns).items():                    # datasets.py:634-636 (mutates shared cfg)
            setattr(self.g, k, v)

    def _crop(self, arr, box, dtype=torch.float32):
        (x1, y1, z1), (x2, y2, z2) = box
        a = np.asarray(arr)[x1:x2, y1:y2, z1:z2]
        if dtype == torch.int32:
            return torch.squeeze(torch.tensor(a.astype(np.float64).astype(int), dtype=torch.int32))
        return torch.squeeze(torch.tensor(a.astype(float), dtype=dtype))

    # -- random setup (datasets.py:466-493) --------------------------------------------------
    def setup(self):
        g, log = self.g, self.log
        if g.low_res_only:
            photo = False
        elif g.left_hemis_only:
            photo = True
        else:
            u = np.random.rand(); log.append(("setup.photo", u))
            photo = u < g.photo_prob
        u = np.random.rand(); log.append(("setup.pathol", u)); pathol = u < g.pathology_prob
        u = np.random.rand(); log.append(("setup.rshape", u)); rshape = u < g.random_shape_prob
        spac = None
        if photo:
            u = np.random.rand(); log.append(("setup.spac", u)); spac = 2.5 + 10 * u
        if g.left_hemis_only:
            flip = False
        else:
            n = np.random.randn(); log.append(("setup.flip", n)); flip = n < g.flip_prob
        if photo:
            res = np.array([self.res[0], spac, self.res[2]])
            thick = np.array([self.res[0], 0.1, self.res[2]])
        else:
            res, thick = draw_resolution(g.low_res_only, log)
        return dict(resolution=res, thickness=thick, photo_mode=photo, pathol_mode=pathol,
                    pathol_random_shape=rshape, spac=spac, flip=flip)

    # -- deformation (datasets.py:187-249) ----------------------------------------------------
    def deformation(self, setups, shp):
        g, log = self.g, self.log
        u = np.random.rand(3); log.append(("aff.rot", u))
        rot = (2 * g.max_rotation * u - g.max_rotation) / 180.0 * np.pi
        u = np.random.rand(3); log.append(("aff.shear", u))
        sh = 2 * g.max_shear * u - g.max_shear
        u = np.random.rand(3); log.append(("aff.scale", u))
        sc = 1 + (2 * g.max_scaling * u - g.max_scaling)
        sfd = np.prod(sc) ** .33333333333
        A = torch.tensor(affine_matrix(rot, sh, sc), dtype=torch.float32)
        c2 = torch.tensor((np.array(shp[:3]) - 1) / 2, dtype=torch.float32)
        if g.random_shift:
            ms = torch.tensor(np.array(shp[:3]) - self.size, dtype=torch.float32) / 2
            ms = ms.clamp(min=0)
            u = torch.rand(3, dtype=torch.float64); log.append(("aff.shift", u.clone()))
            c2 = c2 + (2 * (ms * u) - ms)
        Fld = Fneg = None
        if g.nonlinear_transform:
            u = np.random.rand(1); log.append(("nl.scale", u))
            ns = g.nonlin_scale_min + u * (g.nonlin_scale_max - g.nonlin_scale_min)
            small = np.round(ns * np.array(self.size)).astype(int).tolist()
            if setups["photo_mode"]:
                small[1] = int(np.round(self.size[1] / setups["spac"]))
            u = np.random.rand(); log.append(("nl.std", u))
            std = g.nonlin_std_max * u
            eps = torch.randn([*small, 3], dtype=torch.float32); log.append(("nl.field", eps.clone()))
            Fsmall = std * eps
            Fld = zoom_linear(Fsmall, np.array(self.size) / small)
            if setups["photo_mode"]:
                Fld[..., 1] = 0
            if "surface" in self.tasks:
                n = g.n_steps_svf_integration
                pos, neg = svf_integrate(Fld, self.grid, n), svf_integrate(-Fld, self.grid, n)
                Fld, Fneg = pos, neg
        rel, lo, hi = deform_grid(self.centred, shp, A, c2, Fld)
        return dict(scaling_factor_distances=sfd, A=A, c2=c2, F=Fld, Fneg=Fneg, rel=rel, lo=lo, hi=hi)

    # -- targets (Generator/utils.py:296-477) -------------------------------------------------
    def _warp(self, arr, D, default_max=False, mean=0.0, scale=1.0):
        I = torch.nan_to_num(self._crop(arr, (D["lo"], D["hi"])))
        I -= mean
        I /= scale
        if self.hemis_mask is not None:                  # utils.py:310-311
            I[self.hemis_mask == 0] = 0
        dv = torch.max(I) if default_max else 0.0
        return sample_trilinear(I, *D["rel"], default=dv)

    def target_image(self, key, setups, D):
        I = self._warp(self.vol[key], D)
        I -= torch.min(I)
        I /= torch.max(I)
        if setups["flip"]:
            I = torch.flip(I, [0])
        return {key: I[None]}

    def target_ct(self, setups, D):
        I = self._warp(self.vol["CT"], D, scale=1000)
        if setups["flip"]:
            I = torch.flip(I, [0])
        return {"CT": I[None]}

    def target_distance(self, setups, D):
        if self.hemis_mask is not None:                  # left hemisphere only: two maps (utils.py:373-374)
            lp, lw = [self._warp(v, D, default_max=True, mean=128.0, scale=20) for v in self.vol["distance"][:2]]
            I = torch.stack([lp, lw], 0)
        else:
            lp, lw, rp, rw = [self._warp(v, D, default_max=True, mean=128.0, scale=20) for v in
                              self.vol["distance"]]
            if setups["flip"]:
                lp, rp = torch.flip(rp, [0]), torch.flip(lp, [0])
                lw, rw = torch.flip(rw, [0]), torch.flip(lw, [0])
            I = torch.stack([lp, lw, rp, rw], 0)
        I /= D["scaling_factor_distances"]
        m = self.cfg.max_surf_distance
        return {"distance": torch.clamp(I, min=-m, max=m)}

    def target_registration(self, setups, D):
        r = [self._warp(v, D, scale=10000) for v in self.vol["registration"]]
        if setups["flip"]:
            r = [-torch.flip(r[0], [0]), torch.flip(r[1], [0]), torch.flip(r[2], [0])]
        return {"registration": torch.stack(r, 0)}

    def target_bias_field(self, setups, D):
        I = self._warp(self.vol["bias_field"], D)
        if setups["flip"]:
            I = torch.flip(I, [0])
        return {"bias_field": I[None]}

    def target_segmentation(self, setups, D):
        S = self._crop(self.vol["segmentation"], (D["lo"], D["hi"]), dtype=torch.int32)
        if self.hemis_mask is not None:                  # utils.py:400-401
            S[self.hemis_mask == 0] = 0
        if self.g.deform_one_hots:
            oh = sample_trilinear(self.eye[self.lut[S.long()]], *D["rel"])
        else:
            Sdef = sample_nearest(S, *D["rel"])
            oh = self.eye[self.lut[Sdef.long()]]
        if setups["flip"]:
            oh = torch.flip(oh, [0])[..., self.vflip]
        return {"segmentation": oh.permute(3, 0, 1, 2)}

    def target_pathology(self, setups):
        """read_and_deform_pathology (Generator/utils.py:428-455) for the branches the reference can run: no
        pathology (file_name None) and a random Perlin shape (no advection: datasets.py:603-604)."""
        z = torch.zeros(self.size)[None]
        if not setups["pathol_mode"]:
            return {"pathology": z, "pathology_prob": z.clone()}
        if not setups["pathol_random_shape"]:
            raise NotImplementedError("file-based pathology maps: site-specific paths, and the reference's "
                                      "read_and_deform call omits a required argument (utils.py:442)")
        from oracle import shapeid_oracle as so
        sg = self.cfg.pathology_shape_generator
        u = np.random.random_sample(); self.log.append(("pathol.percentile", u))       # np.random.uniform(a, b)
        percentile = sg.mask_percentile_min + (sg.mask_percentile_max - sg.mask_percentile_min) * u
        res = [int(r) for r in sg.perlin_res]
        shp = (res[0] + 1, res[1] + 1, res[2] + 1)
        th = np.random.rand(*shp); self.log.append(("perlin.theta", th.copy()))
        ph = np.random.rand(*shp); self.log.append(("perlin.phi", ph.copy()))
        theta, phi = 2 * np.pi * th, 2 * np.pi * ph
        g = np.stack((np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta), np.cos(phi)), axis=3)
        g[-1] = g[0]                                                                   # tileable=(True, False, False)
        noise = so.perlin(tuple(self.size), res, g)
        _, prob = so.shape_from_noise(noise, percentile)
        Pdef = torch.from_numpy(prob)
        thr = sg.pathol_thres * Pdef.max()
        P = Pdef.clone()
        P[Pdef < thr] = 0.
        P[Pdef >= thr] = 1.
        if P.mean() <= sg.pathol_tol:
            return {"pathology": z, "pathology_prob": z.clone()}
        return {"pathology": P[None], "pathology_prob": Pdef[None]}

    def encode_pathology(self, I, P, Pprob, direction):
        """datasets.py:496-518 (direction is always given on the synthetic path)."""
        P, Pprob = torch.squeeze(P), torch.squeeze(Pprob)
        I_mu = (I * P).sum() / P.sum()
        p_mask = torch.round(P).long()
        r1 = torch.rand(10000, dtype=torch.float); self.log.append(("pathol.mus", r1.clone()))
        pth_mus = 3 * I_mu / 4 + I_mu / 4 * r1
        pth_mus = pth_mus if direction else -pth_mus
        r2 = torch.rand(10000, dtype=torch.float); self.log.append(("pathol.sigmas", r2.clone()))
        pth_sigmas = I_mu / 4 * r2
        eps = torch.randn(p_mask.shape, dtype=torch.float); self.log.append(("pathol.eps", eps.clone()))
        I += Pprob * (pth_mus[p_mask] + pth_sigmas[p_mask] * eps)
        I[I < 0] = 0
        return I

    def targets(self, setups, D):
        T = {"name": self.case_name}
        for key in ("T1", "T2", "FLAIR"):
            T.update(self.target_image(key, setups, D) if key in self.vol else {key: 0.0})
        for task in self.tasks:
            if task in ("T1", "T2", "FLAIR"):
                continue
            if task == "pathology":
                T.update(self.target_pathology(setups))
                continue
            if task == "surface":            # registered, but never in `modalities` (datasets.py:621-631)
                T["surface"] = 0.0
                continue
            fn = {"CT": self.target_ct, "segmentation": self.target_segmentation,
                  "distance": self.target_distance, "registration": self.target_registration,
                  "bias_field": self.target_bias_field}.get(task)
            if fn is None:
                continue
            T.update(fn(setups, D) if task in self.vol else {task: 0.0})
        return T

    # -- contrast + GMM (datasets.py:357-372, 430-464) -----------------------------------------
    def contrast(self, photo):
        g, log = self.g, self.log
        mu = torch.rand(256, dtype=torch.float32); log.append(("gmm.mu", mu.clone()))
        sg = torch.rand(256, dtype=torch.float32); log.append(("gmm.sigma", sg.clone()))
        mu = 25 + 200 * mu
        sg = 5 + 20 * sg
        u = np.random.rand(); log.append(("gmm.ct", u))
        if u < g.ct_prob:
            for name, (base, span) in (("darker", (25, 10)), ("dark", (90, 20)), ("bright", (110, 20)),
                                       ("brighter", (150, 50))):
                w = torch.rand(1, dtype=torch.float32); log.append(("gmm.ct." + name, w.clone()))
                v = base + span * w[0]
                for l in CT_GROUPS[name]:
                    mu[l] = v
        bg = True
        if not photo:
            u = np.random.rand(1); log.append(("gmm.bg", u)); bg = bool(u < 0.5)
        if bg:
            mu[0] = 0
        v = 0.02 * torch.arange(50)
        for base, (a, b) in ((100, (1, 2)), (150, (2, 3)), (200, (3, 4))):
            mu[base:base + 50] = mu[a] * (1 - v) + mu[b] * v
            sg[base:base + 50] = torch.sqrt(sg[a] ** 2 * (1 - v) + sg[b] ** 2 * v)
        mu[250] = mu[4]
        sg[250] = sg[4]
        return mu, sg

    def synth(self, setups, D, target):
        mu, sg = self.contrast(setups["photo_mode"])
        G = self._crop(self.vol["Gen"], (D["lo"], D["hi"]))
        G[G == 77] = 2
        if self.hemis_mask is not None:                  # datasets.py:367-368
            G[self.hemis_mask == 0] = 0
        Gr = torch.round(G).long()
        eps = torch.randn(Gr.shape, dtype=torch.float32); self.log.append(("gmm.eps", eps.clone()))
        S = mu[Gr] + sg[Gr] * eps
        S[S < 0] = 0
        S = sample_trilinear(S, *D["rel"])
        u = np.random.rand(); self.log.append(("mix.u", u))
        if u < self.cfg.mix_synth_prob:
            v = torch.rand(4); self.log.append(("mix.v", v.clone()))
            v[2] = 0 if "T2" not in self.vol else v[2]
            v[3] = 0 if "FLAIR" not in self.vol else v[3]
            v /= torch.sum(v)
            S = v[0] * S + v[1] * target["T1"][0]
            if "T2" in self.vol:
                S += v[2] * target["T2"][0]
            if "FLAIR" in self.vol:
                S += v[3] * target["FLAIR"][0]
        direction = None
        if isinstance(target.get("pathology"), torch.Tensor) and target["pathology"].sum() > 0:
            # datasets.py:390-406, quirks included (masks of the crop's shape applied to the deformed image;
            # the cerebral image is warped a second time)
            Sc = S.clone()
            Sc[Gr == 0] = 0
            Sc = sample_trilinear(Sc, *D["rel"])[None]
            wm = (Gr == 2) | (Gr == 41)
            wm_mean = (S * wm).sum() / wm.sum()
            gm = (Gr != 0) & (Gr != 2) & (Gr != 41)
            gm_mean = (S * gm).sum() / gm.sum()
            target["pathology"][Sc == 0] = 0
            target["pathology_prob"][Sc == 0] = 0
            direction = bool(gm_mean > wm_mean)
        else:
            target["pathology"] = 0.0
            target["pathology_prob"] = 0.0
        S[S < 0.] = 0.
        return self.augment(S, setups, "synth", target, direction)

    def real(self, input_mode, setups, D, target):
        """Real-image input (augment_sample, datasets.py:306-320): raw crop (no nan_to_num, no clamp at 0),
        hemisphere mask, trilinear warp, CT window; pathology direction fixed by the modality (datasets.py:520-528)."""
        I = self._crop(self.vol[input_mode], (D["lo"], D["hi"]))
        if self.hemis_mask is not None:
            I[self.hemis_mask == 0] = 0
        I = sample_trilinear(I, *D["rel"])
        if input_mode == "CT":
            I = torch.clamp(I, min=0., max=80.)
        direction = input_mode in ("T2", "FLAIR")
        return self.augment(I, setups, input_mode, target, direction)

    # -- augmentation chain (datasets.py:306-354, utils.py:568-638) ----------------------------
    def op_gamma(self, I, aux, setups):
        n = np.random.randn(1)[0]; self.log.append(("gamma.n", n))
        gamma = torch.tensor(np.exp(self.g.gamma_std * n), dtype=torch.float64)
        return 300.0 * (I / 300.0) ** gamma

    def op_bias_field(self, I, aux, setups):
        g = self.g
        if getattr(self, "_mode", "synth") == "CT":      # utils.py:575-577: no bias field on CT
            aux["high_res"] = I
            return I
        u = np.random.rand(1); self.log.append(("bf.scale", u))
        s = g.bf_scale_min + u * (g.bf_scale_max - g.bf_scale_min)
        small = np.round(s * np.array(self.size)).astype(int).tolist()
        if setups["photo_mode"]:
            small[1] = int(np.round(self.size[1] / setups["spac"]))
        u = np.random.rand(1); self.log.append(("bf.std", u))
        std = torch.tensor(g.bf_std_min + (g.bf_std_max - g.bf_std_min) * u, dtype=torch.float32)
        eps = torch.randn(small, dtype=torch.float32); self.log.append(("bf.field", eps.clone()))
        lowres = std * eps
        bflog = zoom_linear(lowres, np.array(self.size) / small)
        out = I * torch.exp(bflog)
        aux["BFlog"] = bflog
        aux["high_res"] = out
        return out

    def op_resample(self, I, aux, setups):
        u = np.random.rand(); self.log.append(("rs.u", u))
        stds = (0.85 + 0.3 * u) * np.log(5) / np.pi * setups["thickness"] / self.res
        stds[setups["thickness"] <= self.res] = 0.0
        B = blur3d(I, stds)
        new = (np.array(self.size) * self.res / setups["resolution"]).astype(int)
        fac = new / np.array(self.size)
        v = [zoom_tables(self.size[d], fac[d], int(new[d]), dtype64=True) for d in range(3)]
        II, JJ, KK = torch.meshgrid(*v, indexing="ij")
        aux["factors"] = fac
        aux["stds"] = stds
        return sample_trilinear(B, II, JJ, KK)

    def op_noise(self, I, aux, setups):
        g = self.g
        u = np.random.rand(1); self.log.append(("noise.u", u))
        sd = torch.tensor(g.noise_std_min + (g.noise_std_max - g.noise_std_min) * u, dtype=torch.float32)
        eps = torch.randn(I.shape, dtype=torch.float32); self.log.append(("noise.eps", eps.clone()))
        out = I + sd * eps
        out[out < 0] = 0
        return out

    def augment(self, I, setups, input_mode, target=None, direction=None):
        if target is not None:
            if isinstance(target.get("pathology"), torch.Tensor) and target["pathology"].sum() > 0:
                I = self.encode_pathology(I, target["pathology"], target["pathology_prob"], direction)
                I[I < 0.] = 0.
            else:
                target["pathology"] = 0.0
                target["pathology_prob"] = 0.0
        aux = {}
        self._mode = input_mode
        steps = self.aug_steps["synth"] if input_mode == "synth" else self.aug_steps["real"]
        table = {"gamma": self.op_gamma, "bias_field": self.op_bias_field, "resample": self.op_resample,
                 "noise": self.op_noise}
        stages = {}
        for name in steps:
            I = table[name](I, aux, setups)
            stages[name] = I
        if getattr(self.g, "bspline_zooming", False):
            I = bspline_resize(I, self.size)
        else:
            I = zoom_linear(I, 1 / aux["factors"])
        top = torch.max(I)
        out = I / top
        flip = setups["flip"]
        sample = {}
        if "super_resolution" in self.tasks:
            r = aux["high_res"] / top - out
            sample["high_res_residual"] = torch.flip(r, [0])[None] if flip else r[None]
        sample["input"] = torch.flip(out, [0])[None] if flip else out[None]
        if "bias_field" in self.tasks and input_mode != "CT":
            b = aux["BFlog"]
            sample["bias_field_log"] = torch.flip(b, [0])[None] if flip else b[None]
        self.stages = stages
        self.aux = aux
        return sample

    # -- __getitem__ (datasets.py:638-681, 700-757) --------------------------------------------
    def sample(self):
        self.log = []
        u = np.random.rand(); self.log.append(("input.mode", u))     # read_input (datasets.py:563-588)
        probs = vars(getattr(self.cfg.modality_probs, self.dataset_name))
        input_mode = "synth"
        for m in ("T1", "T2", "FLAIR", "CT"):
            if u < probs[m] and m in self.vol:
                input_mode = m
                break
        self.input_mode = input_mode
        shp = np.asarray(self.vol["Gen" if input_mode == "synth" else input_mode]).shape
        setups = self.setup()
        D = self.deformation(setups, shp)
        self.hemis_mask = None
        if getattr(self.g, "left_hemis_only", False):    # get_left_hemis_mask (datasets.py:251-262)
            S = self._crop(self.vol["segmentation"], (D["lo"], D["hi"]), dtype=torch.int32)
            S = self.lut[S.long()]
            (x1, y1, z1), (x2, y2, z2) = D["lo"], D["hi"]
            X = torch.squeeze(torch.from_numpy(np.asarray(self.vol["registration"][0])[x1:x2, y1:y2, z1:z2]
                                               .astype(np.float64)))
            self.hemis_mask = ((S > 0) & (X < 0)).int()
        target = self.targets(setups, D)
        def one():
            if input_mode == "synth":
                self._merge(self.cfg.synth_image_generator)
                return self.synth(setups, D, target)
            self._merge(self.cfg.real_image_generator)           # datasets.py:666-669, 737-745
            return self.real(input_mode, setups, D, target)
        if not self.brain_id:
            sample = one()
        else:
            sample = []
            for i in range(self.g.all_samples):
                self._merge(self.cfg.mild_generator if i < self.g.mild_samples else self.cfg.severe_generator)
                sample.append(one())
        if not isinstance(target.get("pathology"), torch.Tensor):
            target["pathology"] = 0.0
            target["pathology_prob"] = 0.0
        elif setups["flip"]:                                                       # datasets.py:672-674
            target["pathology"] = torch.flip(target["pathology"], [1])
            target["pathology_prob"] = torch.flip(target["pathology_prob"], [1])
        self.setups, self.deform = setups, D
        return 1, self.dataset_name, input_mode, target, sample


def seed_all(s):
    np.random.seed(s)
    pyrandom.seed(s)
    torch.manual_seed(s)
