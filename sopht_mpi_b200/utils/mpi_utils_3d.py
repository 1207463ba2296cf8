"""3D topology / halo / field / Lagrangian communicators with the reference's class
names (``sopht_mpi/utils/mpi_utils_3d.py``)."""
import numpy as np

from .comm import (MPIConstruct, MPIFieldCommunicator, MPIGhostCommunicator,
                   MPILagrangianFieldCommunicator)


class MPIConstruct3D(MPIConstruct):
    def __init__(self, grid_size_z, grid_size_y, grid_size_x, periodic_domain=False,
                 real_t=np.float64, rank_distribution=None):
        super().__init__(3, (grid_size_z, grid_size_y, grid_size_x), periodic_domain, real_t,
                         rank_distribution)


class MPIGhostCommunicator3D(MPIGhostCommunicator):
    pass


class MPIFieldCommunicator3D(MPIFieldCommunicator):
    pass


class MPILagrangianFieldCommunicator3D(MPILagrangianFieldCommunicator):
    pass
