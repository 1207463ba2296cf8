"""ctypes binding of ``libsophtb200.so`` (C ABI declared in ``include/sopht_b200.h``).

The product path fails loudly when the CUDA library is missing; there is no CPU
fallback.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_int, c_int32, c_int64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsophtb200.so")
# developer knob: an alternate build of the same library (A/B runs of kernel variants)
LIB_PATH = os.environ.get("SB200_LIB", LIB_PATH)

F32, F64 = 0, 1


class Grid(Structure):
    _fields_ = [
        ("dim", c_int32),
        ("dtype", c_int32),
        ("gs", c_int32),
        ("n", c_int32 * 3),
        ("phys", c_int32 * 6),
    ]


class IBParams(Structure):
    _fields_ = [
        ("lag_dtype", c_int32),
        ("kernel_type", c_int32),
        ("width", c_int32),
        ("substart_xyz", c_int32 * 3),
        ("dx", c_double),
        ("coord_shift", c_double),
        ("stiffness", c_double),
        ("damping", c_double),
    ]


_G = POINTER(Grid)
_P = POINTER(IBParams)
_V = c_void_p

# name -> (restype, argtypes); must list every symbol of include/sopht_b200.h
PROTOTYPES = {
    "sb200_last_error": (c_char_p, []),
    "sb200_version": (c_int, []),
    "sb200_set_fixed_val": (c_int, [c_int, _V, c_int64, c_double, _V]),
    "sb200_elementwise_sum": (c_int, [c_int, _V, _V, _V, c_int64, _V]),
    "sb200_elementwise_copy": (c_int, [c_int, _V, _V, c_int64, _V]),
    "sb200_elementwise_saxpby": (c_int, [c_int, _V, _V, c_double, _V, c_double, c_int64, _V]),
    "sb200_elementwise_cross_product": (c_int, [c_int, _V, _V, _V, c_int64, _V]),
    "sb200_add_fixed_val": (c_int, [c_int, _V, c_int, c_int64, POINTER(c_double), _V]),
    "sb200_update_vorticity_from_velocity_forcing": (c_int, [_G, _V, _V, c_double, _V]),
    "sb200_curl": (c_int, [_G, _V, _V, c_double, _V]),
    "sb200_diffusion_flux": (c_int, [_G, _V, _V, c_double, _V]),
    "sb200_tile_flag_count": (c_int64, [_G]),
    "sb200_update_vorticity_from_sparse_forcing": (c_int, [_G, _V, _V, c_double, _V, _V]),
    "sb200_clear_flagged_tiles": (c_int, [_G, _V, c_int, _V, _V]),
    "sb200_diffusion_timestep": (c_int, [_G, _V, c_int, _V, c_double, _V]),
    "sb200_advection_flux_eno3": (c_int, [_G, _V, _V, _V, c_double, _V]),
    "sb200_advection_timestep_eno3": (c_int, [_G, _V, c_int, _V, _V, c_double, _V]),
    "sb200_divergence": (c_int, [_G, _V, _V, c_double, _V]),
    "sb200_laplacian_filter": (c_int, [_G, _V, c_int, c_int, c_int, _V, _V, _V]),
    "sb200_laplacian_filter_order1_out_of_place": (c_int, [_G, _V, _V, c_int, _V, _V, _V]),
    "sb200_laplacian_filter_axis": (c_int, [_G, _V, _V, c_int, _V]),
    "sb200_laplacian_filter_stage": (c_int, [_G, _V, _V, c_int, _V, c_int, _V]),
    "sb200_clear_physical_ring": (c_int, [_G, _V, c_int, c_int, _V]),
    "sb200_penalise_field_boundary": (c_int, [_G, _V, c_int, c_int, _V, _V]),
    "sb200_brinkmann_penalise": (c_int, [c_int, _V, c_double, _V, _V, _V, c_int, c_int64, _V]),
    "sb200_char_func_from_level_set": (c_int, [c_int, _V, _V, c_double, c_int64, _V]),
    "sb200_update_vorticity_from_penalised_velocity": (c_int, [_G, _V, _V, _V, c_double, _V]),
    "sb200_max_abs_sum": (c_int, [_G, _V, c_int, _V, _V]),
    "sb200_max": (c_int, [_G, _V, c_int, _V, _V]),
    "sb200_sum_squares": (c_int, [_G, _V, c_int, _V, _V]),
    "sb200_velocity_from_stream_function": (
        c_int, [_G, _V, _V, c_double, POINTER(c_double), _V, _V, _V]),
    "sb200_vorticity_rhs_fused_3d": (c_int, [_G, _V, _V, _V, _V, c_double, c_double, _V]),
    "sb200_vorticity_rhs_fused_3d_range": (c_int, [_G, _V, _V, _V, c_double, c_double, c_int, c_int, _V]),
    "sb200_poisson_create": (
        c_int, [POINTER(_V), c_int, c_int, c_int, c_int, c_int, c_int, c_double, c_int, c_int, c_int, _V]),
    "sb200_poisson_destroy": (c_int, [_V]),
    "sb200_poisson_solve": (c_int, [_V, _V, _V, c_int, _V]),
    "sb200_poisson_workspace_bytes": (c_int64, [_V]),
    "sb200_poisson_fft_available": (c_int, []),
    "sb200_enable_peer_access": (c_int, [c_int, c_int]),
    "sb200_peer_copy": (c_int, [_V, c_int, _V, c_int, c_int64, _V]),
    "sb200_peer_copy_blocks": (c_int, [c_int, _V, _V, _V, c_int, c_int64, _V]),
    "sb200_peer_push_blocks": (c_int, [c_int, _V, _V, c_int64, c_int, _V, c_int, _V, _V]),
    "sb200_peer_wait_flags": (c_int, [_V, c_int, c_int, _V, _V]),
    "sb200_peer_alloc": (c_int, [c_int64, POINTER(_V)]),
    "sb200_peer_free": (c_int, [_V]),
    "sb200_ipc_export": (c_int, [_V, _V]),
    "sb200_ipc_open": (c_int, [_V, POINTER(_V)]),
    "sb200_ipc_close": (c_int, [_V]),
    "sb200_poisson_set_profiling": (c_int, [_V, c_int]),
    "sb200_poisson_last_stage_ms": (c_int, [_V, _V, c_int]),
    "sb200_poisson_slab_buffer_bytes": (c_int64, [_V, c_int]),
    "sb200_poisson_slab_forward": (c_int, [_V, _V, c_int, _V, _V]),
    "sb200_poisson_slab_spectral": (c_int, [_V, _V, c_int, _V]),
    "sb200_poisson_slab_backward": (c_int, [_V, _V, c_int, _V, _V]),
    "sb200_ib_interact_lag": (c_int, [_G, _P, c_int64, _V, _V, _V, _V, _V, _V, _V, _V, _V, _V]),
    "sb200_ib_spread": (c_int, [_G, _P, c_int64, _V, _V, _V, _V]),
    "sb200_ib_rank_address": (c_int, [c_int, c_int, c_int64, _V, c_double, POINTER(c_double), POINTER(c_int32), _V,
                                      _V, _V]),
    "sb200_ib_interact_owned": (c_int, [_G, _P, c_int64, _V, _V, _V, _V, _V, _V, _V, _V, _V, _V, c_int, _V]),
    "sb200_ib_spread_owned": (c_int, [_G, _P, c_int64, _V, _V, _V, _V, c_int, _V]),
    "sb200_clear_ghost_cells": (c_int, [_G, _V, c_int, _V]),
    "sb200_ghost_sum_add_z": (c_int, [_G, _V, c_int, _V, _V, _V]),
    "sb200_ib_interpolate": (c_int, [_G, _P, c_int64, c_int, _V, _V, _V, _V]),
    "sb200_ib_update_position_mismatch": (c_int, [c_int, _V, _V, c_int64, c_double, _V]),
}


class SophtB200Error(RuntimeError):
    pass


def bind(path):
    """dlopen ``path`` and attach the prototypes; raises if a symbol is missing."""
    lib = ctypes.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def load():
    """The CUDA library (built by ``python -m sopht_mpi_b200.build``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SophtB200Error(
                f"{LIB_PATH} not found: build it with `python -m sopht_mpi_b200.build` "
                "(there is no CPU fallback)")
        _lib = bind(LIB_PATH)
    return _lib


def check(lib, err):
    if err != 0:
        msg = lib.sb200_last_error()
        raise SophtB200Error(f"libsophtb200 error {err}: {msg.decode() if msg else ''}")


def dtype_code(real_t):
    dt = np.dtype(real_t)
    if dt == np.float32:
        return F32
    if dt == np.float64:
        return F64
    raise ValueError(f"unsupported dtype {dt}")


def make_grid(dim, real_t, gs, local_n, phys):
    """local_n: local interior size in array order ((z,)y,x); phys: 2*dim flags in
    array order (z_prev,z_next,)y_prev,y_next,x_prev,x_next."""
    g = Grid()
    g.dim = dim
    g.dtype = dtype_code(real_t)
    g.gs = gs
    n = [1, 1, 1]
    n[3 - dim:] = [int(v) for v in local_n]
    ph = [1] * 6
    ph[6 - 2 * dim:] = [int(bool(v)) for v in phys]
    for i in range(3):
        g.n[i] = n[i]
    for i in range(6):
        g.phys[i] = ph[i]
    return g
