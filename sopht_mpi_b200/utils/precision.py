"""Precision helpers.

Stand-ins for ``sopht.utils.precision`` (un-vendored dependency of the reference,
used e.g. at reference ``sopht_mpi/simulator/flow/flow_simulators_mpi_3d.py:19,430``).
"""
import numpy as np


def get_real_t(precision: str = "single"):
    """Return the numpy floating type for a precision name."""
    if precision == "single":
        return np.float32
    if precision == "double":
        return np.float64
    raise ValueError("Precision argument must be single or double")


def get_test_tol(precision: str = "single") -> float:
    """Small tolerance that also enters the stable-dt formula
    (reference ``flow_simulators_mpi_3d.py:430-446``): 10 * machine eps."""
    if precision == "single":
        return float(10 * np.finfo(np.float32).eps)
    if precision == "double":
        return float(10 * np.finfo(np.float64).eps)
    raise ValueError("Precision argument must be single or double")
