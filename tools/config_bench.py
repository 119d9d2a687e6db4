"""BASELINE.json configs[2] and configs[3] on one B200 (parity-test configurations, not bench lines; numbers go to
DESIGN.md / profiles):

  interpol : 4-channel 256^3 volume, 7-step scaling-and-squaring of a 3-channel displacement (trilinear, dct2),
             cubic prefilter, cubic grid_pull of the 4 channels, nearest grid_pull of a label volume
             (316*N algorithmic bytes; reference on 8 CPU threads: 63.4 s)
  shapeid  : 192^3 Perlin shape + curl velocity + dopri5 advection, nt = 10
             (reference on 8 CPU threads: 95.6 s, 224 RHS evaluations)

    python tools/config_bench.py [interpol] [shapeid]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best, out


def interpol_cfg(n=256, oracle_check=False):
    from brainfm_b200 import interpol
    torch.manual_seed(0)
    vol = torch.rand(1, 4, n, n, n, device="cuda")
    lab = torch.randint(0, 57, (1, 1, n, n, n), device="cuda").float()
    svf = 2 * torch.randn(1, n, n, n, 3, device="cuda")
    stages = {}

    def ss():
        disp = svf / 2 ** 7
        for _ in range(7):
            grid = interpol.add_identity_grid(disp)
            disp = disp + interpol.grid_pull(disp.permute(0, 4, 1, 2, 3), grid, interpolation=1, bound='dct2',
                                             extrapolate=True).permute(0, 2, 3, 4, 1)
        return disp
    stages["scaling_and_squaring_7_three_call_form"], disp_ref = timed(ss)
    stages_fused, disp = timed(lambda: interpol.exp_velocity(svf, steps=7, bound='dct2', extrapolate=True))
    assert torch.equal(disp, disp_ref), "fused scaling and squaring differs from the composed form"
    del stages["scaling_and_squaring_7_three_call_form"]
    stages["scaling_and_squaring_7"] = stages_fused
    grid = interpol.add_identity_grid(disp)
    stages["cubic_prefilter_4ch"], coeff = timed(lambda: interpol.spline_coeff_nd(vol, interpolation=3, bound='dct2', dim=3))
    stages["cubic_pull_4ch"], out = timed(lambda: interpol.grid_pull(coeff, grid, interpolation=3, bound='dct2',
                                                                     extrapolate=True))
    stages["nearest_pull_labels"], lo = timed(lambda: interpol.grid_pull(lab, grid, interpolation=0, bound='dct2',
                                                                         extrapolate=True))
    total = sum(stages.values())
    N = n ** 3
    # parity spot check at full size: strided sample of the cubic pull and of the label pull against the numpy oracle
    parity = None
    if oracle_check:        # tests only (tests/test_configs_gpu.py): bench.py never touches oracle/ on this path
        from oracle import interpol_oracle as io_
        sel = torch.arange(0, N, 40009, device="cuda")[:256]
        g = grid.reshape(1, N, 3)[:, sel].cpu().numpy()[:, :, None, None, :].astype(np.float64)
        want = io_.pull(coeff.cpu().numpy().astype(np.float64), g, [3, 3, 3], [3, 3, 3], 1)[0, :, :, 0, 0]
        got = out.reshape(4, N)[:, sel].cpu().numpy()
        wl = io_.pull(lab.cpu().numpy().astype(np.float64), g, [0, 0, 0], [3, 3, 3], 1)[0, 0, :, 0, 0]
        parity = {"points": int(sel.numel()), "cubic_max_abs_err": float(np.abs(got - want).max()),
                  "labels_equal": bool(np.array_equal(lo.reshape(N)[sel].cpu().numpy(), wl))}
    return {"config": "interpol 4ch %d^3 (configs[2])" % n, "ms": stages, "total_ms": total,
            "algorithmic_GB": 316 * N / 1e9, "algorithmic_GBps": 316 * N / total / 1e6,
            "oracle_spot_check": parity, "reference_cpu_s_8_threads": 63.4 if n == 256 else None}


def shapeid_cfg(n=192):
    from brainfm_b200.ShapeID import perlin3d as P
    from brainfm_b200.ShapeID.DiffEqs import odeint_adjoint
    from brainfm_b200.ShapeID.DiffEqs.pde import AdvDiffPDE
    shape, res, dt, nt = (n, n, n), [2, 2, 2], 0.1, 10
    np.random.seed(0)
    stages = {}
    for rep in range(2):                     # second pass: warm (the first one pays CUDA module loads)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mask, prob = P.generate_shape_3d(shape, res, 92, 'cuda')
        torch.cuda.synchronize()
        stages["generate_shape_3d"] = 1e3 * (time.perf_counter() - t0)
        t0 = time.perf_counter()
        V = P.generate_velocity_3d(shape, res, 500, 'cuda')
        torch.cuda.synchronize()
        stages["generate_velocity_3d"] = 1e3 * (time.perf_counter() - t0)
    pde = AdvDiffPDE(data_spacing=[1., 1., 1.], perf_pattern='adv', V_type='vector_div_free', V_dict=V, BC='neumann',
                     dt=dt, device='cuda')
    t = torch.from_numpy(np.arange(nt) * dt).cuda()
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sol, solver = odeint_adjoint(pde, prob[None], t, dt, method='dopri5', return_solver=True)
        torch.cuda.synchronize()
        stages["dopri5_nt10"] = 1e3 * (time.perf_counter() - t0)
    N = n ** 3
    S = len(solver.trace)
    alg = (84 * N + S * 368 * N + (nt - 1) * 60 * N)
    return {"config": "ShapeID %d^3 (configs[3])" % n, "ms": stages, "total_ms": sum(stages.values()),
            "rhs_evaluations": int(solver.n_rhs), "steps": S, "algorithmic_GB": alg / 1e9,
            "algorithmic_GBps": alg / sum(stages.values()) / 1e6,
            "mass_drift": float((sol[-1].double().sum() / sol[0].double().sum() - 1).abs()),
            "reference_cpu_s_8_threads": 95.6 if n == 192 else None}


def brainid_cfg(n_items=2, steps=100):
    """configs[4] (a): BrainIDGen stream -- one deformation and one set of targets per item, all_samples = 4
    contrasts (2 mild + 2 severe, demo_synth.yaml:100-101) of 160^3 each; 2 items = 8 samples per step."""
    import tempfile
    import bench
    from brainfm_b200 import io as bio
    from brainfm_b200.Generator import BrainIDGen
    dev = torch.device("cuda", 0)
    subs = bench.make_inputs(n_items)
    root = tempfile.mkdtemp(prefix="bfm_brainid_")
    names = []
    for s, v in enumerate(subs):
        stem = os.path.join(root, "HCP.sub%02d." % s)
        bio.register_volume(stem + "T1w.nii", v["T1"])
        bio.register_volume(stem + "generation_labels.nii", v["Gen"])
        names.append(stem + "T1w.nii")
    with open(os.path.join(root, "train.txt"), "w") as f:
        f.write("\n".join(names) + "\n")
    cfg = bench.bench_cfg()
    cfg.split_root = root
    cfg.dataset_option = "brain_id"
    cfg.generator.all_samples, cfg.generator.mild_samples = 4, 2
    ds = BrainIDGen(cfg, dev)
    ds.write_bflog = True
    np.random.seed(7)
    idxs = list(range(n_items))
    for _ in range(5):
        items = ds.generate_batch(idxs)
    assert ds._native is not None and len(items[0][4]) == 4
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ds.generate_batch(idxs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"config": "BrainIDGen stream 160^3, all_samples 4 (configs[4]a)", "items_per_step": n_items,
            "samples_per_step": 4 * n_items, "ms_per_step": ms,
            "samples_per_s": 4 * n_items / ms * 1e3, "items_per_s": n_items / ms * 1e3,
            "planner": "native"}


def realmix_cfg(steps=100):
    """The bench configuration with the reference's default input table for HCP-like data: a real T1 input with
    probability 0.5, a synthetic one otherwise (cfgs/generator/default.yaml:14-22), planned natively."""
    import bench
    dev = torch.device("cuda", 0)
    ds = bench.build_dataset(bench.make_inputs(bench.BATCH), dev)
    ds.gen_args.modality_probs.HCP.T1 = 0.5
    ds.input_prob = vars(ds.gen_args.modality_probs)
    np.random.seed(3)
    idxs = list(range(bench.BATCH))
    n_real = 0
    for _ in range(5):
        ds.generate_batch(idxs)
    assert ds._native is not None
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        items = ds.generate_batch(idxs)
        n_real += sum(1 for it in items if it[2] != 'synth')
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"config": "BaseGen batch 8x160^3, real T1 input with probability 0.5", "ms_per_step": ms,
            "samples_per_s": bench.BATCH / ms * 1e3, "real_fraction": n_real / (steps * bench.BATCH),
            "planner": "native"}


if __name__ == "__main__":
    which = sys.argv[1:] or ["interpol", "shapeid", "brainid", "realmix"]
    if "realmix" in which:
        print(json.dumps(realmix_cfg()))
    if "brainid" in which:
        print(json.dumps(brainid_cfg()))
    if "interpol" in which:
        print(json.dumps(interpol_cfg(int(os.environ.get("INTERPOL_N", "256")))))
    if "shapeid" in which:
        print(json.dumps(shapeid_cfg(int(os.environ.get("SHAPEID_N", "192")))))
