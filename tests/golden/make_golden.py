"""Generate golden vectors from the REFERENCE's own numba kernels.

Runs only in the build container (needs /root/reference); the resulting small
``.npz`` fixtures are committed and are what pins ``oracle/ib.py``.

    python tests/golden/make_golden.py

The reference modules are loaded by file path; the un-vendored / absent
imports they make at module top (``sopht.utils.field``, ``mpi4py``,
``sopht_mpi.utils.mpi_logger``) are replaced by minimal stand-ins that the
called functions never touch.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference/sopht_mpi"
OUT = os.path.dirname(os.path.abspath(__file__))


def _stub_modules():
    class VectorField:
        @staticmethod
        def x_axis_idx():
            return 0

        @staticmethod
        def y_axis_idx():
            return 1

        @staticmethod
        def z_axis_idx():
            return 2

    for name in ("sopht", "sopht.utils", "sopht.utils.field", "mpi4py",
                 "sopht_mpi", "sopht_mpi.utils", "sopht_mpi.utils.mpi_logger",
                 "sopht_mpi.utils.lab_cmap", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["sopht.utils.field"].VectorField = VectorField
    sys.modules["mpi4py"].MPI = types.SimpleNamespace(PROC_NULL=-1, FLOAT=None, DOUBLE=None)
    sys.modules["sopht_mpi.utils.lab_cmap"].lab_cmap = None
    sys.modules["sopht_mpi.utils.mpi_logger"].logger = types.SimpleNamespace(
        error=print, warning=print, debug=print, info=print)


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_ib(dim, eul_t, lag_t, substart_xyz, tag):
    rng = np.random.default_rng(1000 + dim)
    mod = _load(
        f"{REF}/numeric/immersed_boundary_ops/EulerianLagrangianGridCommunicatorMPI{dim}D.py",
        f"ref_ib{dim}d")
    n_local = 12
    gs, width = 2, 2
    n_lag = 9
    dx = eul_t(1.0 / 32)
    shift = eul_t(dx / 2)
    substart = np.array(substart_xyz)
    # positions inside the local block, some exactly on cell centres/faces
    lo = substart * float(dx)
    pos = (lo.reshape(dim, 1) + rng.uniform(0.0, n_local * float(dx), size=(dim, n_lag))).astype(lag_t)
    pos[:, 0] = (lo + float(dx) * 3).astype(lag_t)           # on a cell face
    pos[:, 1] = (lo + float(dx) * 3.5).astype(lag_t)         # on a cell centre
    pos[:, 2] = (lo + 1e-7).astype(lag_t)                    # at the block start
    shape = (n_local + 2 * gs,) * dim
    eul_vec = rng.uniform(size=(dim,) + shape).astype(eul_t)
    lag_vec = rng.uniform(-1, 1, size=(dim, n_lag)).astype(lag_t)
    sub_shift = substart - gs
    sfx = f"_{dim}d"
    support_k = getattr(mod, f"generate_local_eulerian_grid_support_of_lagrangian_grid_kernel{sfx}")(
        dx=dx, eul_grid_coord_shift=shift, interp_kernel_width=width,
        mpi_local_substart_coord_shift=sub_shift)
    e2l_k = getattr(mod, f"generate_eulerian_to_lagrangian_grid_interpolation_kernel{sfx}")(
        dx=dx, interp_kernel_width=width, n_components=dim)
    e2l_s = getattr(mod, f"generate_eulerian_to_lagrangian_grid_interpolation_kernel{sfx}")(
        dx=dx, interp_kernel_width=width, n_components=1)
    l2e_k = getattr(mod, f"generate_lagrangian_to_eulerian_grid_interpolation_kernel{sfx}")(
        interp_kernel_width=width, n_components=dim)
    cos_k = getattr(mod, f"generate_cosine_interpolation_weights_kernel{sfx}")(
        dx=dx, interp_kernel_width=width, real_t=eul_t)
    pes_k = getattr(mod, f"generate_peskin_interpolation_weights_kernel{sfx}")(
        dx=dx, interp_kernel_width=width, real_t=eul_t)

    kshape = (2 * width,) * dim
    nearest = np.empty((dim, n_lag), dtype=int)
    support = np.empty((dim,) + kshape + (n_lag,), dtype=lag_t)
    support_k(support, nearest, pos)
    support_raw = support.copy()
    w_cos = np.empty(kshape + (n_lag,), dtype=lag_t)
    cos_k(w_cos, support.copy())
    w_pes = np.empty(kshape + (n_lag,), dtype=lag_t)
    pes_k(w_pes, support.copy())
    lag_out = np.zeros((dim, n_lag), dtype=lag_t)
    e2l_k(lag_out, eul_vec, w_cos, nearest)
    lag_out_s = np.zeros((n_lag,), dtype=lag_t)
    e2l_s(lag_out_s, eul_vec[0], w_cos, nearest)
    eul_out = np.zeros((dim,) + shape, dtype=eul_t)
    l2e_k(eul_out, lag_vec, w_cos, nearest)
    np.savez_compressed(
        os.path.join(OUT, f"ib_{tag}.npz"),
        dim=dim, gs=gs, width=width, dx=dx, shift=shift, substart_xyz=substart,
        pos=pos, eul_vec=eul_vec, lag_vec=lag_vec,
        nearest=nearest.astype(np.int64), support=support_raw,
        w_cos=w_cos, w_pes=w_pes, e2l_vec=lag_out, e2l_scalar=lag_out_s, l2e_vec=eul_out)
    print("wrote", tag, "nearest[:, :3] =", nearest[:, :3].tolist())


def make_ownership():
    mod3 = _load(f"{REF}/utils/mpi_utils_3d.py", "ref_mpi_utils_3d")
    mod2 = _load(f"{REF}/utils/mpi_utils_2d.py", "ref_mpi_utils_2d")
    rng = np.random.default_rng(7)
    out = {}
    for lag_t, tname in ((np.float64, "f64"), (np.float32, "f32")):
        # 3D: global 32x16x64 (z,y,x), topology (4,2,1)
        topo = np.array([4, 2, 1])
        gsz = np.array([32, 16, 64])
        local = gsz // topo
        dx = np.float32(1.0 / 64)
        shift = np.float32(dx / 2)
        n = 64
        pos = np.stack([
            rng.uniform(0.0, 64 * float(dx) - 1e-3, n),
            rng.uniform(0.0, 16 * float(dx) - 1e-3, n),
            rng.uniform(0.0, 32 * float(dx) - 1e-3, n),
        ]).astype(lag_t)
        # points exactly on block boundaries (cell-centre-shifted and unshifted)
        pos[2, 0] = lag_t(8 * float(dx))
        pos[2, 1] = lag_t(8 * float(dx) + float(shift))
        pos[2, 2] = lag_t(16 * float(dx) + float(shift))
        pos[1, 3] = lag_t(8 * float(dx) + float(shift))
        pos[:, 4] = lag_t(float(shift) * 0.5)  # below the shift: negative -> trunc to 0
        rank_map = np.arange(int(np.prod(topo)), dtype=np.int32).reshape(topo)
        fake = types.SimpleNamespace(
            eul_grid_coord_shift=shift, eul_subblock_dx=dx * local, rank_map=rank_map,
            mpi_construct=types.SimpleNamespace(grid_topology=topo, grid=None))
        addr = mod3.MPILagrangianFieldCommunicator3D._compute_lag_nodes_rank_address(fake, pos)
        out[f"pos3_{tname}"] = pos
        out[f"addr3_{tname}"] = addr.astype(np.int32)
        # 2D: global 32x64 (y,x), topology (4,1)
        topo2 = np.array([4, 1])
        local2 = np.array([32, 64]) // topo2
        pos2 = pos[:2].copy()
        pos2[1] = pos[2]
        rank_map2 = np.arange(4, dtype=np.int32).reshape(topo2)
        fake2 = types.SimpleNamespace(
            eul_grid_coord_shift=shift, eul_subblock_dx=dx * local2, rank_map=rank_map2,
            mpi_construct=types.SimpleNamespace(grid_topology=topo2, grid=None))
        addr2 = mod2.MPILagrangianFieldCommunicator2D._compute_lag_nodes_rank_address(fake2, pos2)
        out[f"pos2_{tname}"] = pos2
        out[f"addr2_{tname}"] = addr2.astype(np.int32)
    out.update(dx=dx, shift=shift, topo3=topo, local3=local, topo2=topo2, local2=local2)
    np.savez_compressed(os.path.join(OUT, "ownership.npz"), **out)
    print("wrote ownership")


if __name__ == "__main__":
    _stub_modules()
    make_ib(3, np.float32, np.float64, (0, 0, 0), "3d_f32_f64")
    make_ib(3, np.float64, np.float64, (0, 0, 12), "3d_f64_f64_sub")
    make_ib(3, np.float32, np.float32, (0, 0, 24), "3d_f32_f32_sub")
    make_ib(2, np.float64, np.float64, (0, 12), "2d_f64_f64_sub")
    make_ib(2, np.float32, np.float64, (0, 0), "2d_f32_f64")
    make_ownership()
