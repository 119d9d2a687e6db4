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
    """Forget the in-memory volumes AND their device copies (a path may be registered again with new data)."""
    _REGISTRY.clear()
    for c in _SHARED.values():
        c.clear()


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
    padding, which is what the fused gather kernel (bfm_gen_warp, real-image targets) requires.

    Eviction.  The fused chain stores raw device pointers of cached volumes in its descriptors, so a volume must
    stay alive until the kernels of the batch that looked it up have been ENQUEUED (after that, stream order
    protects it: the allocator hands a freed block only to later work).  `begin_batch()` opens an epoch; a volume
    looked up in the current or the previous epoch is never evicted.  When a new volume does not fit under
    `max_bytes`, least-recently-used volumes of older epochs are dropped one at a time BEFORE the new one is
    allocated; if that is not enough, MemoryError is raised (a single batch needs more than the budget) -- the
    cache never silently drops a volume a pending launch still points to.

    Derived tensors (left-hemisphere masks, masked label maps) live in the same cache, are keyed by the paths they
    were computed from and are invalidated when one of those paths is uploaded / refreshed."""

    def __init__(self, device, max_bytes=64 << 30):
        from collections import OrderedDict
        self.device = device
        self.max_bytes = max_bytes
        self._d = OrderedDict()          # (path, kind) -> [tensor, nbytes, epoch]   (LRU order)
        self._derived = {}               # (name, deps) -> [tensor, nbytes]
        self._bytes = 0
        self._epoch = 0
        self.evictions = 0
        self.version = 0                 # bumped whenever the SET of cached tensors changes (insert / evict / clear)

    @staticmethod
    def _pad(shape):
        return int(shape[1]) * int(shape[2]) + int(shape[2]) + 1 if len(shape) >= 3 else 1

    def begin_batch(self):
        """Called by the generator at the start of every batch / item."""
        self._epoch += 1

    @property
    def nbytes(self):
        return self._bytes

    def __contains__(self, key):
        return key in self._d

    def _drop_derived_of(self, path):
        for key in [k for k in self._derived if path in k[1]]:
            self._bytes -= self._derived.pop(key)[1]

    def _make_room(self, nbytes):
        while self._bytes + nbytes > self.max_bytes:
            victim = None
            for key, ent in self._d.items():                 # oldest first
                if ent[2] < self._epoch - 1:
                    victim = key
                    break
            if victim is None:
                if self._derived:                            # derived tensors can always be rebuilt
                    key = next(iter(self._derived))
                    self._bytes -= self._derived.pop(key)[1]
                    continue
                raise MemoryError("DeviceVolumeCache: %d bytes are pinned by the batch in flight, %d more do not fit "
                                  "under max_bytes=%d; raise max_bytes or use smaller batches"
                                  % (self._bytes, nbytes, self.max_bytes))
            self._bytes -= self._d.pop(victim)[1]
            self._drop_derived_of(victim[0])
            self.evictions += 1
            self.version += 1

    def get(self, path, kind="f32"):
        key = (path, kind)
        ent = self._d.get(key)
        if ent is not None:
            ent[2] = self._epoch
            self._d.move_to_end(key)
            return ent[0]
        a = load(path).get_fdata()
        a = np.squeeze(a)
        if kind == "f32":
            h = torch.nan_to_num(torch.from_numpy(np.ascontiguousarray(a.astype(float))).to(torch.float32))
            numel = h.numel() + self._pad(h.shape)
            self._make_room(4 * numel)
            buf = torch.zeros(numel, dtype=torch.float32, device=self.device)
            buf[:h.numel()].copy_(h.reshape(-1))
            t = buf[:h.numel()].view(h.shape)
            nbytes = 4 * numel
        elif kind == "i32":
            h = torch.from_numpy(np.ascontiguousarray(a.astype(int))).to(torch.int32)
            nbytes = 4 * h.numel()
            self._make_room(nbytes)
            t = h.to(self.device)
        elif kind == "gen":
            f = a.astype(np.float32)
            if np.all(f == np.round(f)) and f.min() >= 0 and f.max() <= 255:
                h = torch.from_numpy(np.ascontiguousarray(f.astype(np.uint8)))
            else:
                h = torch.from_numpy(np.ascontiguousarray(f))
            nbytes = h.numel() * h.element_size()
            self._make_room(nbytes)
            t = h.to(self.device)
        else:
            raise ValueError(kind)
        self._d[key] = [t, nbytes, self._epoch]
        self._bytes += nbytes
        self.version += 1
        return t

    def touch(self, keys):
        """Mark cached volumes as used by the current batch (what `get` does on a hit) without returning them."""
        d, epoch = self._d, self._epoch
        for key in keys:
            ent = d[key]
            ent[2] = epoch
            d.move_to_end(key)

    def clear(self):
        self._d.clear()
        self._derived.clear()
        self._bytes = 0
        self.version += 1

    def derived(self, name, deps, build):
        """A tensor computed from cached volumes (`deps`: the paths it depends on), built once by `build()`."""
        key = (name, tuple(deps))
        ent = self._derived.get(key)
        if ent is None:
            t = build()
            ent = [t, t.numel() * t.element_size()]
            self._derived[key] = ent
            self._bytes += ent[1]
        return ent[0]

    def upload(self, path, kind, host_tensor):
        """Overwrite a cached volume with fresh data (pinned host tensor: asynchronous copy on the current stream; or a
        device staging tensor).  The source may be stored in its on-disk integer dtype (uint8 / int16 / int32 ...): the
        conversion to the cached dtype happens on the device (bfm_ingest_volume for 'f32', fused with nan_to_num).
        Derived tensors of the path are invalidated."""
        t = self.get(path, kind)
        self._drop_derived_of(path)
        if kind == "f32" and host_tensor.is_cuda:
            import ctypes as C
            from . import _lib
            code = _INGEST_DTYPES.get(host_tensor.dtype)
            if code is not None and host_tensor.is_contiguous() and host_tensor.numel() == t.numel():
                _lib.check(_lib.lib().bfm_ingest_volume(t.data_ptr(), host_tensor.data_ptr(), code, t.numel(), 1.0, 0.0,
                                                        C.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))))
                return t
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
                                               C.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))))

    def refresh(self, path, kind, host_tensor):
        """upload() + sanitize() on the current stream."""
        t = self.upload(path, kind, host_tensor)
        self.sanitize(path, kind)
        return t


_INGEST_DTYPES = {torch.uint8: 0, torch.int16: 1, torch.int32: 2, torch.float32: 3, torch.int8: 4}

_SHARED = {}


def shared_cache(device):
    """THE volume cache of a device: the generator (inputs, fused targets) and the op-wise target readers
    (read_and_deform & co) look volumes up in the same place, so an upload / refresh is seen by both and a volume
    is held once."""
    key = str(torch.device(device))
    if key not in _SHARED:
        _SHARED[key] = DeviceVolumeCache(torch.device(device))
    return _SHARED[key]
