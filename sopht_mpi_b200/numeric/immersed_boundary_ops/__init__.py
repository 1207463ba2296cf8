from .eulerian_lagrangian_grid_communicator import (
    EulerianLagrangianGridCommunicatorMPI2D,
    EulerianLagrangianGridCommunicatorMPI3D,
    MPIGhostSumCommunicator,
)
from .virtual_boundary_forcing import VirtualBoundaryForcingMPI
