"""Checks shared by the 2D and 3D operator generators."""
import sys


def check_valid_ghost_size_and_kernel_support(ghost_size, kernel_support):
    """A stencil of half-width ``kernel_support`` needs at least that many ghost layers; operator
    generators call this while they are being built, so a bad combination fails before any field
    exists (same contract and ``ValueError`` as the reference's ``sopht_mpi/utils/mpi_utils.py:17-24``)."""
    if ghost_size >= kernel_support:
        return
    generator = sys._getframe(1).f_code.co_name
    raise ValueError(f"{generator}: ghost_size = {ghost_size} cannot hold a stencil of support "
                     f"{kernel_support} (ghost_size >= kernel_support is required)")
