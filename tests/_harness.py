"""Shared helpers of the GPU parity tests: run the oracle on a named case, then feed its recorded draws to
the CUDA generator."""
import os
import tempfile

import numpy as np
import torch

from oracle import gen_oracle as go
from oracle import make_golden as mg


def oracle_case(name):
    (item, orc) = mg.run_oracle(name)
    return item, orc


def cuda_case(name, log, dataset_option=None, run=None, planner='auto'):
    from brainfm_b200 import io as bio
    from brainfm_b200.draws import ReplayDraws
    from brainfm_b200.Generator import dataset_options
    size, src, kind, seed, over, extra, option, stride = mg.CASES[name]
    vols = mg.build_volumes(src, kind, extra)
    root = tempfile.mkdtemp(prefix="bfm_case_")
    stem = os.path.join(root, "HCP.sub01.")
    bio.clear_registry()
    bio.register_volume(stem + "T1w.nii", vols["T1"])
    bio.register_volume(stem + "generation_labels.nii", vols["Gen"])
    bio.register_volume(stem + "brainseg_with_extracerebral.nii", vols["segmentation"])
    if "T2" in vols:
        bio.register_volume(stem + "T2w.nii", vols["T2"])
    if "CT" in vols:
        bio.register_volume(stem + "CT.nii", vols["CT"])
    if "distance" in vols:
        for k, v in zip(["lp_dist_map", "lw_dist_map", "rp_dist_map", "rw_dist_map"], vols["distance"]):
            bio.register_volume(stem + k + ".nii", v)
    if "registration" in vols:
        for k, v in zip(["mni_reg.x", "mni_reg.y", "mni_reg.z"], vols["registration"]):
            bio.register_volume(stem + k + ".nii", v)
    with open(os.path.join(root, "train.txt"), "w") as f:
        f.write(stem + "T1w.nii\n")
    args = mg.cfg_for(size, over, option, ref=False)
    args.split_root = root
    draws = ReplayDraws(log)
    ds = dataset_options[dataset_option or option](args, "cuda", draws=draws, planner=planner)
    item = ds[0] if run is None else run(ds)
    torch.cuda.synchronize()
    return item, ds, draws


def to_np(v):
    return v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
