"""Advection PDE right-hand side (mirror of ShapeID/DiffEqs/pde.py:563-640 for the configuration the generator
uses: perf_pattern 'adv', V_type 'vector_div_free', Neumann / no boundary condition)."""
import torch
import torch.nn as nn

from ... import _lib
from .._common import fvec, ivec, need_cuda, stream
from ..misc import gradient_b, gradient_c, gradient_f  # noqa: F401  (re-exported like the reference module)


class AdvDiffPDE(nn.Module):
    """dC/dt = -(V . grad C) with per-component upwinding; one stencil kernel per evaluation."""

    def __init__(self, data_spacing, perf_pattern, D_type='scalar', V_type='vector', BC=None, dt=0.1, V_dict={},
                 D_dict={}, stochastic=False, device='cpu'):
        super(AdvDiffPDE, self).__init__()
        self.BC = BC
        self.dt = dt
        self.dimension = len(data_spacing)
        self.data_spacing = list(data_spacing)
        self.perf_pattern = perf_pattern
        self.D_type, self.V_type = D_type, V_type
        self.stochastic = stochastic
        self.V_dict, self.D_dict = V_dict, D_dict
        self.Sigma, self.Sigma_V, self.Sigma_D = 0., 0., 0.
        if self.dimension not in (1, 2, 3):
            raise ValueError('Unsupported dimension: %d' % self.dimension)
        self.n_evals = 0

    def _check(self):
        if self.dimension != 3 or 'diff' in self.perf_pattern or self.V_type != 'vector_div_free' or self.stochastic:
            raise NotImplementedError("AdvDiffPDE: only the 3-D 'adv' pattern with V_type='vector_div_free' is built")
        if self.BC not in (None, 'neumann', 'cauchy'):
            raise NotImplementedError('Unsupported B.C.!')

    def forward(self, t, batch_C, out=None):
        """t: scalar; batch_C: (batch, slc, row, col) float32/float64 -> float32 derivative."""
        self._check()
        need_cuda(batch_C, "batch_C")
        C_ = batch_C.contiguous()
        V = self.V_dict
        if out is None:
            out = torch.empty(C_.shape, dtype=torch.float32, device=C_.device)
        L = _lib.lib()
        for b in range(C_.shape[0]):
            _lib.check(L.bfm_advect_rhs(C_[b].data_ptr(), 1 if C_.dtype == torch.float64 else 0,
                                        V['Vx'].data_ptr(), V['Vy'].data_ptr(), V['Vz'].data_ptr(),
                                        ivec(C_.shape[1:]), 1 if self.BC in ('neumann', 'cauchy') else 0,
                                        fvec(self.data_spacing), out[b].data_ptr(), stream()))
        self.n_evals += 1
        return out
