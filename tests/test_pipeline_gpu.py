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
