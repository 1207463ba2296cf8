from .ops import *  # noqa: F401,F403
from .ops import OpContext
from .poisson import UnboundedPoissonSolverMPI2D, UnboundedPoissonSolverMPI3D
