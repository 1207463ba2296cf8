"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the immersed-boundary path.

numpy restatement of the reference's numba kernels
(``sopht_mpi/numeric/immersed_boundary_ops/EulerianLagrangianGridCommunicatorMPI3D.py``
and ``...MPI2D.py``, ``VirtualBoundaryForcingMPI.py``) and of the Lagrangian
rank-ownership map (``sopht_mpi/utils/mpi_utils_3d.py:1352-1384``).
Pinned by golden vectors generated from the reference's own numba kernels:
``tests/golden/make_golden.py`` -> ``tests/golden/ib_*.npz``.

Lagrangian arrays are ``(dim, N)`` in x,y,z order; Eulerian arrays ``(z,y,x)``.
"""
import numpy as np


def support_and_nearest_index(lag_positions, dx, eul_grid_coord_shift, width,
                              substart_xyz, gs):
    """reference ...MPI3D.py:116-178 / ...MPI2D.py twin.

    ``substart_xyz``: local interior start index of this rank (x,y[,z] order).
    Returns (nearest_index int64 (dim,N), support (dim, 2w,..,2w, N))."""
    dim, n = lag_positions.shape
    shift = (np.asarray(substart_xyz) - gs).astype(np.int64)  # mpi_local_substart_coord_shift
    nearest = np.empty((dim, n), dtype=np.int64)
    nearest[...] = (lag_positions - eul_grid_coord_shift) // dx - shift.reshape(dim, 1)
    off = np.arange(-width + 1, width + 1)
    if dim == 3:
        zg, yg, xg = np.meshgrid(off, off, off, indexing="ij")
        idx = np.stack((xg, yg, zg))
    else:
        # reference MPI2D: x_grid, y_grid = np.meshgrid(x, x); stack((x_grid, y_grid))
        xg, yg = np.meshgrid(off, off)
        idx = np.stack((xg, yg))
    kshape = (2 * width,) * dim
    support = (
        (
            nearest.reshape((dim,) + (1,) * dim + (n,))
            + idx.reshape((dim,) + kshape + (1,))
            + shift.reshape((dim,) + (1,) * dim + (1,))
        )
        * dx
        + eul_grid_coord_shift
        - lag_positions.reshape((dim,) + (1,) * dim + (n,))
    )
    return nearest, support.astype(lag_positions.dtype)


def cosine_weights(support, dx, real_t):
    """reference ...MPI3D.py:443-479 (scales support in place by 1/dx first)."""
    dim = support.shape[0]
    support /= dx
    w = real_t((0.25 / dx) ** dim)
    for d in range(dim):
        w = w * (real_t(1.0) + np.cos(real_t(0.5 * np.pi) * support[d]))
    return w.astype(support.dtype)


def peskin_weights(support, dx, real_t):
    """reference ...MPI3D.py:482-589."""
    dim = support.shape[0]
    support[...] = np.fabs(support) / dx
    w = (0.125 / dx) ** dim
    for d in range(dim):
        r = support[d]
        w = w * (
            (r < 1.0) * (3.0 - 2 * r + np.sqrt(np.fabs(1 + 4 * r - 4 * r ** 2)))
            + (r >= 1.0) * (r < 2.0) * (5.0 - 2 * r - np.sqrt(np.fabs(-7 + 12 * r - 4 * r ** 2)))
        )
    return np.asarray(w).astype(real_t).astype(support.dtype)


def _window(nearest, width, i):
    dim = nearest.shape[0]
    # array axes are reversed w.r.t. x,y,z component order
    return tuple(
        slice(nearest[d, i] - width + 1, nearest[d, i] + width + 1)
        for d in reversed(range(dim))
    )


def eulerian_to_lagrangian(lag_field, eul_field, weights, nearest, dx, width):
    """reference ...MPI3D.py:181-326: sum(window * W) * dx**dim.
    lag_field (N,) or (dim,N); eul_field (z,y,x) or (dim,z,y,x)."""
    dim = nearest.shape[0]
    n = nearest.shape[1]
    vec = eul_field.ndim == dim + 1
    for i in range(n):
        win = _window(nearest, width, i)
        if vec:
            for c in range(eul_field.shape[0]):
                lag_field[c, i] = np.sum(eul_field[c][win] * weights[..., i]) * (dx ** dim)
        else:
            lag_field[i] = np.sum(eul_field[win] * weights[..., i]) * (dx ** dim)


def lagrangian_to_eulerian(eul_field, lag_field, weights, nearest, width):
    """reference ...MPI3D.py:329-427: sequential over points, no dx factor."""
    dim = nearest.shape[0]
    n = nearest.shape[1]
    vec = eul_field.ndim == dim + 1
    for i in range(n):
        win = _window(nearest, width, i)
        if vec:
            for c in range(eul_field.shape[0]):
                eul_field[c][win] += (lag_field[c, i] * weights[..., i]).astype(eul_field.dtype)
        else:
            eul_field[win] += (lag_field[i] * weights[..., i]).astype(eul_field.dtype)


def _window_indices(nearest, width):
    """index arrays (array-axis order) of shape (N, 2w, .., 2w): point-major, so that a flattened
    scatter visits the points in the reference's sequential order"""
    dim, n = nearest.shape
    off = np.arange(-width + 1, width + 1)
    idx = []
    for ax in range(dim):  # array axis ax <-> component dim - 1 - ax
        shape = [1] * (dim + 1)
        shape[1 + ax] = 2 * width
        idx.append(nearest[dim - 1 - ax].reshape((n,) + (1,) * dim) + off.reshape(shape))
    return tuple(np.broadcast_arrays(*idx))


def eulerian_to_lagrangian_fast(lag_field, eul_field, weights, nearest, dx, width):
    """Vectorised twin of :func:`eulerian_to_lagrangian` (gathers all windows at once) for the
    timed CPU baseline; same products, numpy's pairwise summation instead of the loop's."""
    dim = nearest.shape[0]
    idx = _window_indices(nearest, width)
    w = np.moveaxis(weights, -1, 0)  # (N, 2w, .., 2w)
    axes = tuple(range(1, dim + 1))
    if eul_field.ndim == dim + 1:
        for c in range(eul_field.shape[0]):
            lag_field[c] = np.sum(eul_field[c][idx] * w, axis=axes) * (dx ** dim)
    else:
        lag_field[...] = np.sum(eul_field[idx] * w, axis=axes) * (dx ** dim)


def lagrangian_to_eulerian_fast(eul_field, lag_field, weights, nearest, width):
    """Vectorised twin of :func:`lagrangian_to_eulerian`: ``np.add.at`` applies the updates in
    index order, i.e. point after point like the reference loop."""
    dim = nearest.shape[0]
    idx = _window_indices(nearest, width)
    w = np.moveaxis(weights, -1, 0)
    bshape = (-1,) + (1,) * dim
    if eul_field.ndim == dim + 1:
        for c in range(eul_field.shape[0]):
            np.add.at(eul_field[c], idx, (lag_field[c].reshape(bshape) * w).astype(eul_field.dtype))
    else:
        np.add.at(eul_field, idx, (lag_field.reshape(bshape) * w).astype(eul_field.dtype))


def clear_ghost_cells_nd(field, gs, dim):
    """reference ...MPI3D.py:786-792 (last ``dim`` axes)."""
    for ax in range(-1, -dim - 1, -1):
        sl = [slice(None)] * field.ndim
        sl[ax] = slice(0, gs)
        field[tuple(sl)] = 0
        sl[ax] = slice(-gs, None)
        field[tuple(sl)] = 0


def lag_nodes_rank_address(global_lag_positions, dx, eul_grid_coord_shift,
                           local_grid_size, grid_topology):
    """reference mpi_utils_3d.py:1352-1384 / mpi_utils_2d.py:611-637.
    ``local_grid_size`` and ``grid_topology`` in array order (z,y,x)/(y,x);
    positions in x,y[,z] order.  Truncation toward zero, dtype of positions.
    Cartesian rank = row-major index of the block coordinates (MPI
    ``Create_cart(reorder=False)``)."""
    dim = global_lag_positions.shape[0]
    sub_dx = dx * np.asarray(local_grid_size)
    coords = []
    for ax in range(dim):  # array axis order
        pos = global_lag_positions[dim - 1 - ax]
        coords.append(((pos - eul_grid_coord_shift) / sub_dx[ax]).astype(np.int32))
    for ax in range(dim):
        if np.any(coords[ax] >= grid_topology[ax]):
            raise RuntimeError("Lagrangian node is found outside of Eulerian domain!")
    rank_map = np.arange(int(np.prod(grid_topology)), dtype=np.int32).reshape(grid_topology)
    return rank_map[tuple(coords)]


class VirtualBoundaryForcingOracle:
    """Single-rank restatement of reference VirtualBoundaryForcingMPI.py:333-459
    (all points local)."""

    def __init__(self, k, c, grid_dim, dx, real_t, lag_dtype, gs,
                 eul_grid_coord_shift=None, width=2, kernel_type="cosine", fast=False):
        self.k, self.c = k, c
        # fast: vectorised gather / scatter (the reference's kernels are numba-compiled loops; a
        # python-level loop over 1e5 points would misrepresent its speed in the CPU baseline)
        self._e2l = eulerian_to_lagrangian_fast if fast else eulerian_to_lagrangian
        self._l2e = lagrangian_to_eulerian_fast if fast else lagrangian_to_eulerian
        self.dim = grid_dim
        self.dx = dx
        self.real_t = real_t
        self.gs = gs
        self.width = width
        self.kernel_type = kernel_type
        self.shift = real_t(dx / 2) if eul_grid_coord_shift is None else eul_grid_coord_shift
        self.lag_dtype = lag_dtype
        self.time = 0.0
        self.position_mismatch = None
        self.velocity_mismatch = None

    def _ensure(self, n):
        if self.position_mismatch is None or self.position_mismatch.shape[1] != n:
            self.position_mismatch = np.zeros((self.dim, n), dtype=self.lag_dtype)
            self.velocity_mismatch = np.zeros_like(self.position_mismatch)
        self.flow_velocity = np.zeros((self.dim, n), dtype=self.lag_dtype)
        self.forcing = np.zeros((self.dim, n), dtype=self.lag_dtype)

    def compute_interaction_force_on_lag_grid(self, eul_velocity, lag_pos, lag_vel):
        n = lag_pos.shape[1]
        self._ensure(n)
        self.nearest, support = support_and_nearest_index(
            lag_pos, self.dx, self.shift, self.width, (0,) * self.dim, self.gs
        )
        if self.kernel_type == "cosine":
            self.weights = cosine_weights(support, self.dx, self.real_t)
        else:
            self.weights = peskin_weights(support, self.dx, self.real_t)
        self._e2l(self.flow_velocity, eul_velocity, self.weights, self.nearest, self.dx, self.width)
        self.velocity_mismatch[...] = self.flow_velocity - lag_vel
        self.forcing[...] = self.k * self.position_mismatch + self.c * self.velocity_mismatch

    def compute_interaction_force_on_eul_and_lag_grid(self, eul_forcing, eul_velocity,
                                                      lag_pos, lag_vel):
        self.compute_interaction_force_on_lag_grid(eul_velocity, lag_pos, lag_vel)
        self._l2e(eul_forcing, self.forcing, self.weights, self.nearest, self.width)
        # single rank: every neighbour is PROC_NULL -> ghost sum only clears ghosts
        clear_ghost_cells_nd(eul_forcing, self.gs, self.dim)

    def time_step(self, dt):
        self.position_mismatch[...] = self.position_mismatch + dt * self.velocity_mismatch
        self.time += dt
