"""Generic checks shared by 2D and 3D operators (reference ``sopht_mpi/utils/mpi_utils.py``)."""
import inspect


def _get_caller_name(steps=1):
    frame = inspect.currentframe().f_back
    for _ in range(steps):
        frame = frame.f_back
    return frame.f_code.co_name


def check_valid_ghost_size_and_kernel_support(ghost_size, kernel_support):
    """reference ``mpi_utils.py:17-24``"""
    if ghost_size < kernel_support:
        raise ValueError(
            f"Inconsistent ghost_size={ghost_size} and kernel_support="
            f"{kernel_support} for kernel {_get_caller_name(steps=1)}. "
            "Need to have ghost_size >= kernel_support"
        )
