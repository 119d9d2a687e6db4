"""Deterministic synthetic inputs shared by the golden-fixture generator, the tests and bench.py.

Nothing here reads /root/reference: the label maps are procedural so that
they exist on the GPU box too.
"""
import types

import numpy as np


def block_labels(shape, block=8, n_labels=57):
    """label[i,j,k] = (i//b + 3*(j//b) + 7*(k//b)) % n_labels  (SURVEY.md 8d)."""
    i, j, k = np.meshgrid(*[np.arange(n) for n in shape], indexing="ij", sparse=True)
    return ((i // block + 3 * (j // block) + 7 * (k // block)) % n_labels).astype(np.float32)


def brain_like_labels(shape, seed=7):
    """Nested ellipsoid shells with label values drawn from the generation-label range
    (0..99 tissue classes plus a few partial-volume labels 100..250 and the 77 lesion label)."""
    rng = np.random.RandomState(seed)
    i, j, k = np.meshgrid(*[np.linspace(-1, 1, n) for n in shape], indexing="ij", sparse=True)
    r = np.sqrt((i / 0.9) ** 2 + (j / 0.8) ** 2 + (k / 0.85) ** 2)
    wob = 0.05 * np.sin(7 * i) * np.cos(5 * j) + 0.05 * np.sin(6 * k)
    shell = np.clip(((r + wob) * 12).astype(int), 0, 15)
    table = np.array([2, 77, 3, 41, 4, 17, 120, 42, 175, 8, 230, 24, 250, 85, 0, 0], dtype=np.float32)
    lab = table[shell]
    speck = rng.rand(*shape) < 0.002
    lab[speck & (r < 0.7)] = 77
    return lab.astype(np.float32)


def smooth_image(shape, phase=0.0):
    """A smooth positive 'real modality' volume."""
    i, j, k = np.meshgrid(*[np.arange(n, dtype=np.float64) for n in shape], indexing="ij", sparse=True)
    v = 120 + 60 * np.sin(i / 7.0 + phase) * np.cos(j / 9.0) + 40 * np.sin(k / 5.0 - phase) + 0.1 * (i + j + k)
    return v.astype(np.float32)


def seg_labels(shape):
    """Integer segmentation map whose values are members of the 56-label brainseg list."""
    return (block_labels(shape, block=6, n_labels=57)).astype(np.float32)


def ns(**kw):
    return types.SimpleNamespace(**kw)


def default_cfg(size=(160, 160, 160), **gen_over):
    """The Namespace tree default.yaml + train/brain_id.yaml produce (cfgs/generator/default.yaml:57-123,
    cfgs/generator/train/brain_id.yaml:50-126), written out literally so it exists without the reference."""
    gen = dict(size=list(size), left_hemis_only=False, low_res_only=False, photo_prob=0.2, max_rotation=15,
               max_shear=0.2, max_scaling=0.2, nonlin_scale_min=0.03, nonlin_scale_max=0.06, nonlin_std_max=4,
               bag_prob=0.5, bag_scale_min=0.02, bag_scale_max=0.08, bf_scale_min=0.02, bf_scale_max=0.04,
               bf_std_min=0.1, bf_std_max=0.6, gamma_std=0.1, noise_std_min=0.05, noise_std_max=1.,
               exvixo_prob=0.25, exvixo_prob_vs_photo=0.66666666666666, pv=True, random_shift=False,
               deform_one_hots=False, integrate_deformation_fields=False, produce_surfaces=False,
               bspline_zooming=False, n_steps_svf_integration=8, nonlinear_transform=True, ct_prob=0,
               flip_prob=0.5, pathology_prob=0., random_shape_prob=0., augment_pathology=False,
               mild_samples=0, all_samples=1, all_contrasts=1, num_deformations=1)
    gen.update(gen_over)
    task = dict(T1=False, T2=False, FLAIR=False, CT=False, segmentation=False, distance=False, bias_field=True,
                registration=False, super_resolution=False, age=False, surface=False, pathology=False,
                contrastive=False)
    return ns(
        dataset_names=["HCP"], split="train", split_root=None, data_root=None, dataset_option="default",
        segment_prefix="brainseg_with_extracerebral", mix_synth_prob=0.0, max_surf_distance=3.,
        modality_probs=ns(HCP=ns(T1=0., T2=0., FLAIR=0., CT=0., synth=1.)),
        task=ns(**task),
        augmentation_steps=ns(synth=["gamma", "bias_field", "resample", "noise"],
                              real=["gamma", "bias_field", "resample", "noise"]),
        generator=ns(**gen),
        synth_image_generator=ns(noise_std_min=5., noise_std_max=15.),
        real_image_generator=ns(noise_std_min=0., noise_std_max=0.02),
        mild_generator=ns(bag_prob=0.1, bag_scale_min=0.01, bag_scale_max=0.02, bf_scale_min=0.01,
                          bf_scale_max=0.02, bf_std_min=0., bf_std_max=0.02, gamma_std=0.01, noise_std_min=0.,
                          noise_std_max=0.02),
        severe_generator=ns(bag_prob=0.5, bag_scale_min=0.02, bag_scale_max=0.08, bf_scale_min=0.02,
                            bf_scale_max=0.04, bf_std_min=0.1, bf_std_max=0.6, gamma_std=0.1, noise_std_min=0.05,
                            noise_std_max=1.),
        pathology_shape_generator=ns(perlin_res=[2, 2, 2], mask_percentile_min=85, mask_percentile_max=99.9,
                                     integ_method="dopri5", bc="neumann", V_multiplier=500, dt=0.1, max_nt=10,
                                     pathol_thres=0.5, pathol_tol=0.0000001),
    )
