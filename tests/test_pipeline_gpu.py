"""HostPipeline (host buffers in, host buffers out): per-volume and batched uploads give the same volumes as the
device-resident path; uploads really replace the cached volumes; pinned output slots are reused safely."""
import os
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dataset(n=4, size=48):
    import bench
    from brainfm_b200 import io as bio
    bio.clear_registry()
    old = bench.SIZE
    bench.SIZE = size
    try:
        subs = bench.make_inputs(n)
        ds = bench.build_dataset(subs, torch.device("cuda", 0))
    finally:
        bench.SIZE = old
    return ds, subs


def _run(ds, fn, seed=3):
    np.random.seed(seed)
    ds._native_planner([0, 1, 2, 3])
    ds._native.seed, ds._native.counter = None, 0
    return fn()


def test_batched_and_per_volume_uploads_match_the_resident_path():
    from brainfm_b200.pipeline import HostPipeline
    ds, subs = _dataset()
    idx = [0, 1, 2, 3]
    ref = _run(ds, lambda: torch.cat([it[4]['input'] for it in ds.generate_batch(idx)], 0).cpu().numpy())
    lab_paths = [ds.names[0][s][:-7] + "generation_labels.nii" for s in idx]
    t1_paths = [ds.names[0][s] for s in idx]
    lab = [torch.from_numpy(subs[s]["Gen"].astype(np.uint8)).pin_memory() for s in idx]
    t1 = [torch.from_numpy(subs[s]["T1"]).pin_memory() for s in idx]
    pipe = HostPipeline(ds, depth=3)
    per_volume = [u for s in idx for u in ((lab_paths[s], "gen", lab[s]), (t1_paths[s], "f32", t1[s]))]
    batched = [(lab_paths, "gen", torch.stack(lab).pin_memory()), (t1_paths, "f32", torch.stack(t1).pin_memory())]
    for uploads in (per_volume, batched):
        # poison the cached volumes first: the result can only be right if the upload really lands before the batch
        for p in lab_paths:
            ds.cache.get(p, "gen").zero_()
        for p in t1_paths:
            ds.cache.get(p, "f32").fill_(float("nan"))
        items, host = _run(ds, lambda: pipe.submit(idx, uploads).wait())
        assert host.is_pinned() and tuple(host.shape) == (4, 48, 48, 48)
        assert np.array_equal(host.numpy(), ref[:, ...].reshape(host.shape))
        assert all(torch.isfinite(it[3]['T1']).all() for it in items)


def test_pipeline_in_flight_batches_do_not_clobber_each_other():
    from brainfm_b200.pipeline import HostPipeline
    ds, subs = _dataset()
    idx = [0, 1, 2, 3]
    lab_paths = [ds.names[0][s][:-7] + "generation_labels.nii" for s in idx]
    labs = torch.stack([torch.from_numpy(subs[s]["Gen"].astype(np.uint8)) for s in idx]).pin_memory()
    pipe = HostPipeline(ds, depth=3)
    np.random.seed(9)
    tickets = [pipe.submit(idx, [(lab_paths, "gen", labs)]) for _ in range(3)]
    outs = [t.wait()[1].clone() for t in tickets]
    # different random draws per batch: three different results, each finite and normalised
    assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[2])
    for o in outs:
        assert torch.isfinite(o).all() and float(o.min()) >= 0 and abs(float(o.amax()) - 1) < 1e-6


def test_default_dataloader_collation_like_the_reference_trainer():
    """The reference consumes the dataset through torch's DataLoader with default collation and num_workers 0
    (train.py:133-137; MetricLogger indexes dataset_name[0], input_mode[0]): the 5-tuple must collate."""
    from torch.utils.data import DataLoader
    ds, subs = _dataset()
    np.random.seed(4)
    loader = DataLoader(ds, batch_size=2, shuffle=False, num_workers=0)
    n = 0
    for datasets_num, dataset_name, input_mode, target, sample in loader:
        assert dataset_name[0] == 'HCP' and input_mode[0] == 'synth'
        assert tuple(sample['input'].shape) == (2, 1, 48, 48, 48) and sample['input'].is_cuda
        assert tuple(sample['bias_field_log'].shape) == (2, 1, 48, 48, 48)
        assert tuple(target['T1'].shape) == (2, 1, 48, 48, 48)
        assert len(target['name']) == 2
        n += 1
    assert n == 2 and len(ds) == 4


def test_device_pipeline_two_lanes_equals_one_stream():
    """DevicePipeline (two batches in flight on their own streams and scratch) produces the batches generate_batch
    produces one after the other: same planner stream, same kernels, no cross-talk between the lanes."""
    from brainfm_b200.pipeline import DevicePipeline
    ds, subs = _dataset()
    idx = [0, 1, 2, 3]

    def sequential():
        return [torch.cat([it[4]['input'] for it in ds.generate_batch(idx)], 0).clone() for _ in range(4)]

    def piped():
        pipe = DevicePipeline(ds, depth=2)
        tickets = [pipe.submit(idx) for _ in range(4)]
        outs = [torch.cat([it[4]['input'] for it in t.wait()], 0).clone() for t in tickets]
        t1 = [torch.cat([it[3]['T1'] for it in t.items], 0).clone() for t in tickets]
        return outs, t1

    ref = _run(ds, sequential)
    got, t1 = _run(ds, piped)
    torch.cuda.synchronize()
    for a, b in zip(ref, got):
        assert torch.equal(a, b)
    assert not torch.equal(got[0], got[1])
    for t in t1:
        assert torch.isfinite(t).all() and float(t.min()) == 0.0 and float(t.max()) == 1.0


def test_one_call_batches_equal_the_general_path():
    """generate_batch_fast (bfm_plan_run: one library call per batch, lazily built tuples) against generate_batch with
    the same planner state: every tensor of every item bit for bit, same tuple structure; arbitrary index order."""
    ds, subs = _dataset()
    for idx in ([0, 1, 2, 3], [3, 1], [2, 2, 0]):
        ref = _run(ds, lambda: ds.generate_batch(idx))
        ref = [(a, b, c, {k: (v.clone() if torch.is_tensor(v) else v) for k, v in d.items()},
                {k: v.clone() for k, v in e.items()}) for a, b, c, d, e in ref]
        got = _run(ds, lambda: ds.generate_batch_fast(idx))
        assert got is not None and len(got) == len(idx)
        torch.cuda.synchronize()
        assert got.input.shape[0] == len(idx) and set(got.targets) == {'T1'}
        for n, (r, g) in enumerate(zip(ref, got)):
            assert r[:3] == g[:3]
            assert set(r[4]) == set(g[4]) and {k for k in r[3]} == {k for k in g[3]}
            for k in r[4]:
                assert torch.equal(r[4][k], g[4][k]), (n, k)
            for k, v in r[3].items():
                if torch.is_tensor(v):
                    assert torch.equal(v, g[3][k]), (n, k)
                else:
                    assert v == g[3][k], (n, k)
            assert torch.equal(got.input[n], g[4]['input']) and torch.equal(got.targets['T1'][n], g[3]['T1'])


def test_cache_eviction_is_lru_and_spares_the_batch_in_flight():
    from brainfm_b200 import io as bio
    bio.clear_registry()
    dev = torch.device("cuda", 0)
    vols = {"/v/%d.nii" % k: np.full((16, 16, 16), k, np.float32) for k in range(6)}
    for p, v in vols.items():
        bio.register_volume(p, v)
    one = 4 * (16 ** 3 + 16 * 16 + 16 + 1)
    cache = bio.DeviceVolumeCache(dev, max_bytes=3 * one + 100)
    cache.begin_batch()
    a = [cache.get("/v/%d.nii" % k) for k in range(3)]
    ptrs = [t.data_ptr() for t in a]
    with pytest.raises(MemoryError):                       # a fourth volume in the SAME batch does not fit
        cache.get("/v/3.nii")
    assert cache.evictions == 0 and all(("/v/%d.nii" % k, "f32") in cache for k in range(3))
    cache.begin_batch()                                    # previous batch: still protected (its launches may be pending)
    with pytest.raises(MemoryError):
        cache.get("/v/3.nii")
    cache.begin_batch()
    cache.get("/v/1.nii")                                  # touch 1: volume 0 is now the least recently used
    cache.begin_batch()
    cache.begin_batch()
    t3 = cache.get("/v/3.nii")
    assert cache.evictions == 1 and ("/v/0.nii", "f32") not in cache and ("/v/1.nii", "f32") in cache
    assert float(t3.mean()) == 3.0 and cache.nbytes <= cache.max_bytes
    # the evicted volume comes back from the registry when asked for again
    cache.begin_batch(); cache.begin_batch()
    assert float(cache.get("/v/0.nii").mean()) == 0.0 and cache.evictions == 2
    assert [float(t.mean()) for t in a] == [0.0, 1.0, 2.0] and ptrs == [t.data_ptr() for t in a]   # callers' refs stay valid
    bio.clear_registry()


def test_upload_reaches_every_reader_and_drops_derived_tensors():
    """One cache per device: a refreshed segmentation / image volume is what the op-wise target readers see."""
    from oracle import make_golden as mg
    from tests._harness import cuda_case, oracle_case
    from brainfm_b200.Generator import utils as gu
    name = "g64_full_s4"
    _, orc = oracle_case(name)
    got, ds, _ = cuda_case(name, orc.log)
    assert gu._cache(ds.device) is ds.cache
    seg_path = ds.modalities['segmentation']
    before = got[3]['segmentation'].clone()
    new_seg = torch.zeros(ds.cache.get(seg_path, 'i32').shape, dtype=torch.int32)          # all background
    ds.cache.derived('probe', (seg_path,), lambda: torch.ones(4, device=ds.device))
    ds.cache.upload(seg_path, 'i32', new_seg.pin_memory())
    assert not any(seg_path in k[1] for k in ds.cache._derived)
    ds.rng.pos = 0
    got2 = ds[0]
    torch.cuda.synchronize()
    after = got2[3]['segmentation']
    assert not torch.equal(before, after)
    present = after.amax(dim=(1, 2, 3))                                          # a constant label map: one class
    assert float(present.sum()) == 1.0 and float(after.sum(0).min()) == 1.0


@pytest.mark.parametrize("dtype,code", [(np.uint8, 0), (np.int16, 1), (np.int32, 2), (np.float32, 3), (np.int8, 4)])
def test_ingest_volume_dtypes(dtype, code):
    import ctypes as C
    from brainfm_b200 import _lib
    rng = np.random.RandomState(code)
    n = 160 * 157 + 3                                       # not a multiple of the vector width
    if dtype == np.float32:
        a = rng.randn(n).astype(np.float32) * 100
        a[::97] = np.nan
        a[5::131] = np.inf
        a[7::131] = -np.inf
    else:
        info = np.iinfo(dtype)
        a = rng.randint(info.min, int(info.max) + 1, size=n, dtype=np.int64).astype(dtype)
    src = torch.from_numpy(a).cuda()
    dst = torch.empty(n, dtype=torch.float32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for slope, inter in ((1.0, 0.0), (0.5, -3.0)):
        _lib.check(_lib.lib().bfm_ingest_volume(dst.data_ptr(), src.data_ptr(), code, n, slope, inter, st))
        want = torch.nan_to_num(torch.from_numpy(a.astype(np.float32)) * np.float32(slope) + np.float32(inter))
        assert torch.equal(dst.cpu(), want)
    # unaligned views take the scalar path
    _lib.check(_lib.lib().bfm_ingest_volume(dst.data_ptr() + 4, src.data_ptr() + a.itemsize, code, n - 1, 1.0, 0.0, st))
    assert torch.equal(dst[1:].cpu(), torch.nan_to_num(torch.from_numpy(a[1:].astype(np.float32))))
