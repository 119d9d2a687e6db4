"""Runs the cubic prefilter of configs[2] a few times (ncu target).  Development tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from brainfm_b200 import interpol
vol = torch.rand(1, 4, 256, 256, 256, device="cuda")
for _ in range(2):
    out = interpol.spline_coeff_nd(vol, interpolation=3, bound='dct2', dim=3)
torch.cuda.synchronize()
