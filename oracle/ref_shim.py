"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Import shim that lets the UNMODIFIED reference (`/root/reference`, jhuldr/BrainFM) run in the
build container so that golden fixtures can be generated from it (see oracle/make_golden.py).
`/root/reference` does not exist on the GPU box, so nothing under tests/, smoke() or bench.py
may import this module at run time; only oracle/make_golden.py and the `-m "not gpu"`
tests guarded by `ref_available()` do.

The reference needs a handful of packages that are not installed (nibabel, SimpleITK,
matplotlib, visdom, simplejson, pytz, iopath).  None of them is on the arithmetic path
(nibabel is file I/O, the rest is trainer logging), so they are replaced by empty stub modules;
`nibabel.load` is served from an in-memory table of volumes.
"""
import gzip
import os
import struct
import sys
import tempfile
import types

import numpy as np

REF_ROOT = os.environ.get("BFM_REFERENCE_ROOT", "/root/reference")

_VOLUMES = {}          # path -> np.ndarray served by the fake nibabel.load
_installed = False


def ref_available():
    return os.path.isdir(os.path.join(REF_ROOT, "Generator"))


class _FakeProxy:
    """`img.get_fdata()[a:b, c:d, e:f]` -- float64 like nibabel."""

    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, idx):
        return np.asarray(self._arr[idx], dtype=np.float64)

    def astype(self, *a, **k):
        return np.asarray(self._arr, dtype=np.float64).astype(*a, **k)

    @property
    def shape(self):
        return self._arr.shape


class _FakeImage:
    def __init__(self, arr):
        self._arr = arr
        self.shape = arr.shape
        self.affine = np.eye(4)

    def get_fdata(self):
        return _FakeProxy(self._arr)


def register_volume(path, arr):
    _VOLUMES[path] = np.asarray(arr)


def clear_volumes():
    _VOLUMES.clear()


def _fake_load(path):
    if path in _VOLUMES:
        return _FakeImage(_VOLUMES[path])
    raise FileNotFoundError(path)


def install():
    """Register the stub modules and put the reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not ref_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)

    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    stub("nibabel", load=_fake_load, Nifti1Image=object, save=lambda *a, **k: None)
    stub("SimpleITK")
    mpl = stub("matplotlib", use=lambda *a, **k: None)
    plt = stub("matplotlib.pyplot")
    mpl.pyplot = plt
    stub("visdom", Visdom=object)
    stub("simplejson", dumps=lambda *a, **k: "", loads=lambda *a, **k: {})
    stub("pytz", timezone=lambda *a, **k: None)

    class _PMF:
        @staticmethod
        def get(**kw):
            return None

    io_ = stub("iopath")
    common = stub("iopath.common")
    fio = stub("iopath.common.file_io", PathManagerFactory=_PMF, PathManager=object, g_pathmgr=None)
    io_.common = common
    common.file_io = fio
    try:
        import torchvision  # noqa: F401
    except Exception:
        stub("torchvision")
    try:
        import skimage  # noqa: F401
    except Exception:
        stub("skimage")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _installed = True


def read_mgh(path=None):
    """Minimal MGH/MGZ reader (gzip, 284-byte big-endian header, Fortran order)."""
    path = path or os.path.join(REF_ROOT, "files", "gca.mgz")
    opener = gzip.open if path.endswith("z") else open
    with opener(path, "rb") as f:
        raw = f.read()
    ver, w, h, d, nf, typ, dof = struct.unpack(">7i", raw[:28])
    dt = {0: ">u1", 1: ">i4", 3: ">f4", 4: ">i2"}[typ]
    n = w * h * d * nf
    data = np.frombuffer(raw, dtype=dt, count=n, offset=284)
    shape = (w, h, d) if nf == 1 else (w, h, d, nf)
    return np.ascontiguousarray(data.reshape(shape, order="F").astype(np.float32))


def load_generator_cfg(overrides=None, cfg_file="train/brain_id.yaml"):
    """Namespace tree the reference builds from default.yaml + <cfg_file>."""
    install()
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        import utils.misc as rmisc
        args = rmisc.preprocess_cfg(["cfgs/generator/default.yaml",
                                     os.path.abspath(os.path.join("cfgs/generator", cfg_file))])
    finally:
        os.chdir(cwd)
    for k, v in (overrides or {}).items():
        node = args
        parts = k.split(".")
        for p in parts[:-1]:
            node = getattr(node, p)
        setattr(node, parts[-1], v)
    return args


def make_subject(volumes, name="HCP.sub01"):
    """Register the in-memory volumes of one fake subject; returns (split_root, t1_path).

    `volumes` maps suffixes ('T1w', 'generation_labels', 'brainseg_with_extracerebral',
    'lp_dist_map', ...) to arrays (datasets.py:520-543 naming).
    """
    root = tempfile.mkdtemp(prefix="bfm_oracle_")
    t1 = os.path.join(root, name + ".T1w.nii")
    for suffix, arr in volumes.items():
        register_volume(os.path.join(root, "%s.%s.nii" % (name, suffix)), arr)
    with open(os.path.join(root, "train.txt"), "w") as f:
        f.write(t1 + "\n")
    return root, t1
