"""brainfm_b200.interpol -- CUDA backend with the API of the reference's vendored torch-interpol
(utils/interpol/api.py:3-5, 450-452; resize.py:13; restrict.py:9)."""
from .api import (grid_pull, grid_push, grid_count, grid_grad, spline_coeff, spline_coeff_nd, identity_grid,
                  add_identity_grid, add_identity_grid_, affine_grid, pull, push, count, resize, restrict,
                  compose_step, exp_velocity)
from . import backend

__all__ = ['grid_pull', 'grid_push', 'grid_count', 'grid_grad', 'spline_coeff', 'spline_coeff_nd',
           'identity_grid', 'add_identity_grid', 'add_identity_grid_', 'affine_grid', 'resize', 'restrict']
