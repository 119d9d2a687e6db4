"""Volume access for the generator: an in-memory registry, minimal NIfTI-1 / MGH readers (the reference
uses nibabel, Generator/utils.py:266-270,300-304), and a device-resident cache so that a volume is decoded
and uploaded once instead of once per sample."""
import gzip
import os
import struct

import numpy as np
import torch

_REGISTRY = {}


class Volume:
    """The slice of nibabel's image API the generator touches: .shape, .affine, .get_fdata()."""

    def __init__(self, data, affine=None):
        self._data = data
        self.shape = tuple(data.shape)
        self.affine = np.eye(4) if affine is None else affine

    def get_fdata(self):
        return np.asarray(self._data, dtype=np.float64)

    @property
    def dataobj(self):
        return self._data


def register_volume(path, array, affine=None):
    _REGISTRY[path] = Volume(np.asarray(array), affine)


_SURFACES = {}


def register_surface(path, arrays):
    """In-memory stand-in for a `<case>.mat` surface file: dict with Vlw/Flw/Vrw/Frw/Vlp/Flp/Vrp/Frp."""
    _SURFACES[path] = {k: np.asarray(v) for k, v in arrays.items()}


def load_surface(path):
    """The eight arrays of a surface file (scipy.io.loadmat in the reference, Generator/utils.py:483)."""
    if path in _SURFACES:
        return _SURFACES[path]
    from scipy.io.matlab import loadmat
    return loadmat(path)


def clear_registry():
    _REGISTRY.clear()


def exists(path):
    return path in _REGISTRY or os.path.isfile(path)


def _read_nifti(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rb") as f:
        raw = f.read()
    le = struct.unpack("<i", raw[:4])[0] == 348
    e = "<" if le else ">"
    dim = struct.unpack(e + "8h", raw[40:56])
    datatype = struct.unpack(e + "h", raw[70:72])[0]
    vox_offset = int(struct.unpack(e + "f", raw[108:112])[0])
    slope, inter = struct.unpack(e + "2f", raw[112:120])
    dt = {2: "u1", 4: "i2", 8: "i4", 16: "f4", 64: "f8", 256: "i1", 512: "u2", 768: "u4"}[datatype]
    shape = dim[1:1 + dim[0]]
    data = np.frombuffer(raw, dtype=e + dt, count=int(np.prod(shape)), offset=vox_offset).reshape(shape, order="F")
    if slope not in (0.0, 1.0) or inter != 0.0:
        if slope != 0.0 and not np.isnan(slope):
            data = data * slope + inter
    sform = struct.unpack(e + "h", raw[254:256])[0]
    aff = np.eye(4)
    if sform > 0:
        aff[:3] = np.array(struct.unpack(e + "12f", raw[280:328])).reshape(3, 4)
    else:
        pix = struct.unpack(e + "8f", raw[76:108])
        aff[0, 0], aff[1, 1], aff[2, 2] = pix[1], pix[2], pix[3]
    return Volume(data, aff)


def _read_mgh(path):
    op = gzip.open if path.endswith("z") else open
    with op(path, "rb") as f:
        raw = f.read()
    _, w, h, d, nf, typ, _ = struct.unpack(">7i", raw[:28])
    dt = {0: ">u1", 1: ">i4", 3: ">f4", 4: ">i2"}[typ]
    data = np.frombuffer(raw, dtype=dt, count=w * h * d * nf, offset=284)
    shape = (w, h, d) if nf == 1 else (w, h, d, nf)
    return Volume(data.reshape(shape, order="F"))


def load(path):
    """nib.load replacement: registry first, then NIfTI / MGH files (with the reference's '.gz' retry,
    Generator/utils.py:299-302)."""
    if path in _REGISTRY:
        return _REGISTRY[path]
    for p in (path, path + ".gz"):
        if os.path.isfile(p):
            if p.endswith((".mgz", ".mgh")):
                return _read_mgh(p)
            return _read_nifti(p)
    raise FileNotFoundError(path)


class DeviceVolumeCache:
    """path -> device tensor, decoded once.  kinds: 'f32' (images), 'i32' (segmentation labels),
    'gen' (generation labels: uint8 when every value is an integer in [0,255], else float32).

    'f32' volumes are stored finite (torch.nan_to_num, which the reference applies at every crop read,
    Generator/utils.py:305, is applied once here) and followed by one plane + one row + one voxel of zero
    padding, which is what the fused gather kernel (bfm_gen_warp, real-image targets) requires."""

    def __init__(self, device, max_bytes=64 << 30):
        self.device = device
        self.max_bytes = max_bytes
        self._d = {}
        self._bytes = 0

    @staticmethod
    def _pad(shape):
        return int(shape[1]) * int(shape[2]) + int(shape[2]) + 1 if len(shape) >= 3 else 1

    def get(self, path, kind="f32"):
        key = (path, kind)
        t = self._d.get(key)
        if t is not None:
            return t
        a = load(path).get_fdata()
        a = np.squeeze(a)
        if kind == "f32":
            h = torch.nan_to_num(torch.from_numpy(np.ascontiguousarray(a.astype(float))).to(torch.float32))
            buf = torch.zeros(h.numel() + self._pad(h.shape), dtype=torch.float32, device=self.device)
            buf[:h.numel()].copy_(h.reshape(-1))
            t = buf[:h.numel()].view(h.shape)
        elif kind == "i32":
            t = torch.from_numpy(np.ascontiguousarray(a.astype(int))).to(torch.int32).to(self.device)
        elif kind == "gen":
            f = a.astype(np.float32)
            if np.all(f == np.round(f)) and f.min() >= 0 and f.max() <= 255:
                t = torch.from_numpy(np.ascontiguousarray(f.astype(np.uint8))).to(self.device)
            else:
                t = torch.from_numpy(np.ascontiguousarray(f)).to(self.device)
        else:
            raise ValueError(kind)
        nbytes = t.numel() * t.element_size()
        if self._bytes + nbytes > self.max_bytes:
            self._d.clear()
            self._bytes = 0
        self._d[key] = t
        self._bytes += nbytes
        return t

    def upload(self, path, kind, host_tensor):
        """Overwrite a cached volume with fresh host data: asynchronous copy from pinned memory on the current
        stream.  'f32' volumes must be passed through sanitize() before they are used."""
        t = self.get(path, kind)
        t.copy_(host_tensor, non_blocking=True)
        return t

    def sanitize(self, path, kind="f32"):
        """torch.nan_to_num in place on the current stream (Generator/utils.py:305)."""
        import ctypes as C
        from . import _lib
        if kind != "f32":
            return
        t = self.get(path, kind)
        _lib.check(_lib.lib().bfm_sanitize_f32(t.data_ptr(), t.numel(),
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def refresh(self, path, kind, host_tensor):
        """upload() + sanitize() on the current stream."""
        t = self.upload(path, kind, host_tensor)
        self.sanitize(path, kind)
        return t
