from .precision import get_real_t, get_test_tol
from .field import VectorField
from .logger import logger, RankLogger
from .device import DeviceField
from .comm import MPI
from .mpi_utils import check_valid_ghost_size_and_kernel_support
from .mpi_utils_2d import (MPIConstruct2D, MPIGhostCommunicator2D, MPIFieldCommunicator2D,
                           MPILagrangianFieldCommunicator2D)
from .mpi_utils_3d import (MPIConstruct3D, MPIGhostCommunicator3D, MPIFieldCommunicator3D,
                           MPILagrangianFieldCommunicator3D)
from .mpi_io import MPIIO
