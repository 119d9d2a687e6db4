"""Where the time of configs[3] (ShapeID 192^3: Perlin shape, curl velocity, dopri5 advection) goes: cProfile of
the host side and CUDA-event timing of the individual kernels.  Development tool."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def ev_time(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main(n=192):
    from brainfm_b200.ShapeID import perlin3d as P
    from brainfm_b200.ShapeID.DiffEqs import odeint_adjoint
    import importlib
    O = importlib.import_module('brainfm_b200.ShapeID.DiffEqs.odeint')
    from brainfm_b200.ShapeID.DiffEqs.pde import AdvDiffPDE
    shape, res, dt, nt = (n, n, n), [2, 2, 2], 0.1, 10
    np.random.seed(0)
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mask, prob = P.generate_shape_3d(shape, res, 92, 'cuda')
        torch.cuda.synchronize()
        print("generate_shape_3d rep %d: %.2f ms" % (rep, 1e3 * (time.perf_counter() - t0)))
    noise = P._noise_device(shape, res, (True, False, False), torch.device('cuda'))
    print("perlin kernel: %.3f ms" % ev_time(lambda: P._noise_device(shape, res, (True, False, False), torch.device('cuda'))))
    print("percentile: %.3f ms" % ev_time(lambda: P._percentile_device(noise, 92), 5))
    V = P.generate_velocity_3d(shape, res, 500, 'cuda')
    pde = AdvDiffPDE(data_spacing=[1., 1., 1.], perf_pattern='adv', V_type='vector_div_free', V_dict=V, BC='neumann',
                     dt=dt, device='cuda')
    y = prob[None]
    print("rhs eval (f64 state): %.3f ms" % ev_time(lambda: pde(torch.tensor(0.), y)))
    ks = [pde(torch.tensor(0.), y) for _ in range(7)]
    print("combine 6 stages + y0: %.3f ms" % ev_time(lambda: O._combine(y, ks[:6], [0.1] * 6)))
    print("combine error (7 stages): %.3f ms" % ev_time(lambda: O._combine(None, ks, [0.1] * 7)))
    t = torch.from_numpy(np.arange(nt) * dt).cuda()
    odeint_adjoint(pde, y, t, dt, method='dopri5')
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    sol, solver = odeint_adjoint(pde, y, t, dt, method='dopri5', return_solver=True)
    pr.disable()
    torch.cuda.synchronize()
    print("dopri5: %.2f ms, %d steps, %d rhs" % (1e3 * (time.perf_counter() - t0), len(solver.trace), solver.n_rhs))
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        odeint_adjoint(pde, y, t, dt, method='dopri5')
        torch.cuda.synchronize()
    ev = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    print("GPU busy %.2f ms" % (sum(e.device_time_total for e in ev) * 1e-3))
    for e in ev[:16]:
        print("%9.1f us x%-4d %s" % (e.device_time_total, e.count, e.key[:100]))


if __name__ == "__main__":
    main(int(os.environ.get("SHAPEID_N", "192")))
