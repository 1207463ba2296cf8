"""Shared by the emulation (CPU) and the GPU tests of the Lagrangian ownership kernel."""
import ctypes

import numpy as np

from sopht_mpi_b200 import _lib


def rank_address(call_fn, ptr, pos, dx, shift, local, topo, to_dev=lambda a: a, to_host=lambda a: a):
    """sb200_ib_rank_address with the constants VirtualBoundaryForcingMPI derives (eul_grid_dx *
    local_grid_size as float64, topology in array order)."""
    dim, n = pos.shape
    sub = np.asarray(dx * np.asarray(local), dtype=np.float64)
    pad = 3 - dim
    sub_dx = (ctypes.c_double * 3)(*([1.0] * pad + [float(v) for v in sub]))
    topo_c = (ctypes.c_int32 * 3)(*([1] * pad + [int(v) for v in topo]))
    p_d = to_dev(np.ascontiguousarray(pos))
    out, flag = to_dev(np.full(n, -7, np.int32)), to_dev(np.zeros(1, np.int32))
    call_fn("sb200_ib_rank_address", _lib.dtype_code(pos.dtype), dim, n, ptr(p_d), float(shift), sub_dx, topo_c,
            ptr(out), ptr(flag), None)
    return to_host(out), int(to_host(flag)[0])
