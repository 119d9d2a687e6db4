"""Advection-diffusion PDE right-hand side (mirror of ShapeID/DiffEqs/pde.py:563-640): perf_pattern 'adv' (what the
generator uses), 'diff' and 'adv_diff'; divergence-free vector velocities; constant or scalar diffusivity; Neumann /
no boundary condition; optional stochastic term."""
import torch
import torch.nn as nn

from ... import _lib
from .._common import fvec, ivec, need_cuda, stream
from ..misc import gradient_b, gradient_c, gradient_f  # noqa: F401  (re-exported like the reference module)


class AdvDiffPDE(nn.Module):
    """dC/dt = -(V . grad C) + div(D grad C); one stencil kernel per part and evaluation."""

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

    _V_DIV_FREE = ('vector_div_free', 'vector_div_free_clebsch', 'vector_div_free_stream',
                   'vector_div_free_stream_gauge')

    def _check(self):
        adv, diff = 'diff' not in self.perf_pattern or 'adv' in self.perf_pattern, 'diff' in self.perf_pattern
        if self.dimension != 3:
            raise NotImplementedError("AdvDiffPDE: only the 3-D solver is built")
        if adv and self.V_type not in self._V_DIV_FREE:
            raise NotImplementedError("AdvDiffPDE: advection is built for the divergence-free vector velocities "
                                      "(V_type %r is not)" % self.V_type)
        if diff and self.D_type not in ('constant', 'scalar'):
            raise NotImplementedError("AdvDiffPDE: diffusion is built for D_type 'constant' and 'scalar' "
                                      "(%r is not)" % self.D_type)
        if self.BC not in (None, 'neumann', 'cauchy'):
            raise NotImplementedError('Unsupported B.C.!')

    def forward(self, t, batch_C, out=None):
        """t: scalar; batch_C: (batch, slc, row, col) float32/float64 -> float32 derivative (pde.py:612-640):
        advection -(V . grad C) with per-component upwinding, and / or diffusion div(D grad C) in the reference's
        composition of one-sided differences; `stochastic` adds Sigma * sqrt(dt) * N(0, 1)."""
        self._check()
        need_cuda(batch_C, "batch_C")
        C_ = batch_C.contiguous()
        V = self.V_dict
        if out is None:
            out = torch.empty(C_.shape, dtype=torch.float32, device=C_.device)
        L = _lib.lib()
        adv = 'diff' not in self.perf_pattern or 'adv' in self.perf_pattern
        diff = 'diff' in self.perf_pattern
        neumann = 1 if self.BC in ('neumann', 'cauchy') else 0
        dbl = 1 if C_.dtype == torch.float64 else 0
        Dfield, Dconst = None, 0.0
        if diff:
            D = self.D_dict['D']
            if self.D_type == 'scalar':
                Dfield = torch.as_tensor(D, dtype=torch.float32, device=C_.device).contiguous()
                if Dfield.dim() == 4 and Dfield.shape[0] == 1:
                    Dfield = Dfield[0]
                if tuple(Dfield.shape) != tuple(C_.shape[1:]):
                    raise ValueError("D_dict['D'] must have the spatial shape of the state")
            else:
                Dconst = float(D)
        for b in range(C_.shape[0]):
            if adv:
                _lib.check(L.bfm_advect_rhs(C_[b].data_ptr(), dbl, V['Vx'].data_ptr(), V['Vy'].data_ptr(),
                                            V['Vz'].data_ptr(), ivec(C_.shape[1:]), neumann, fvec(self.data_spacing),
                                            out[b].data_ptr(), stream()))
            if diff:
                _lib.check(L.bfm_diffuse_rhs(C_[b].data_ptr(), dbl, None if Dfield is None else Dfield.data_ptr(), Dconst,
                                             ivec(C_.shape[1:]), neumann, fvec(self.data_spacing), 1 if adv else 0,
                                             out[b].data_ptr(), stream()))
        if self.stochastic:
            import math
            out = out + self.Sigma * math.sqrt(self.dt) * torch.randn_like(C_).to(out)
        self.n_evals += 1
        return out
