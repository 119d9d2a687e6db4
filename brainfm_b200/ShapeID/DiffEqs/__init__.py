from .odeint import odeint, SOLVERS
from .adjoint import odeint_adjoint
from .pde import AdvDiffPDE
