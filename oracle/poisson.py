"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the unbounded Poisson solver.

Single-rank restatement, with ``scipy.fft``, of
``sopht_mpi/numeric/eulerian_grid_ops/poisson_solver_3d/UnboundedPoissonSolverMPI3D.py:22-187``
(+ ``fft_mpi_3d.py:34-48``: forward unnormalised, backward normalised) and of the
2D twin ``poisson_solver_2d/UnboundedPoissonSolverMPI2D.py:12-153``.
The reference's own FFT test pins the transform against ``scipy.fft.rfftn``
(``tests/.../test_poisson_solver_3d/test_fft_mpi_3d.py:93-115``), which is the
library used here.
"""
import numpy as np
import scipy.fft as sfft


class UnboundedPoissonSolverOracle3D:
    def __init__(self, grid_size_z, grid_size_y, grid_size_x, x_range=1.0,
                 real_t=np.float64, workers=1):
        self.nz, self.ny, self.nx = grid_size_z, grid_size_y, grid_size_x
        self.real_t = real_t
        self.workers = workers
        self.x_range = x_range
        self.y_range = x_range * (grid_size_y / grid_size_x)
        self.z_range = x_range * (grid_size_z / grid_size_x)
        self.dx = real_t(x_range / grid_size_x)
        self.greens_function_field = self.construct_greens_function_field()
        ghat = sfft.rfftn(self.greens_function_field, workers=workers)
        # reference :62-64 (complex array times real_t scalar)
        self.fourier_greens_function_times_dx_cubed = ghat * (self.dx ** 3)

    def construct_greens_function_field(self):
        """reference :67-114 on one rank (global_start_idx = 0)."""
        t = self.real_t
        dx = self.dx

        def line(n2):
            return np.linspace(0 * dx, (n2 - 1) * dx, n2).astype(t)

        x = line(2 * self.nx)[None, None, :]
        y = line(2 * self.ny)[None, :, None]
        z = line(2 * self.nz)[:, None, None]
        # the reference builds a meshgrid first (:98-100); broadcasting the three coordinate lines
        # performs the same real_t operations per element without the three 8 N^3 temporaries
        r = np.sqrt(
            (np.minimum(x, 2 * self.x_range - x) ** 2
             + np.minimum(y, 2 * self.y_range - y) ** 2)
            + np.minimum(z, 2 * self.z_range - z) ** 2
        )
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            g = (1 / r) / (4 * np.pi)
        g[0, 0, 0] = 1 / (4 * np.pi * dx)
        return g.astype(t)

    def solve(self, solution_field, rhs_field, gs):
        """reference :133-167; fields are padded local arrays, interiors only."""
        nz, ny, nx = self.nz, self.ny, self.nx
        doubled = np.zeros((2 * nz, 2 * ny, 2 * nx), dtype=self.real_t)
        inner = (slice(gs, -gs),) * 3 if gs > 0 else (slice(None),) * 3
        doubled[:nz, :ny, :nx] = rhs_field[inner]
        fourier = sfft.rfftn(doubled, workers=self.workers)
        conv = fourier * self.fourier_greens_function_times_dx_cubed
        back = sfft.irfftn(conv, s=doubled.shape, workers=self.workers)
        solution_field[inner] = back[:nz, :ny, :nx]

    def vector_field_solve(self, solution_vector_field, rhs_vector_field, gs):
        for c in range(3):
            self.solve(solution_vector_field[c], rhs_vector_field[c], gs)


class UnboundedPoissonSolverOracle2D:
    def __init__(self, grid_size_y, grid_size_x, x_range=1.0, real_t=np.float64, workers=1):
        self.ny, self.nx = grid_size_y, grid_size_x
        self.real_t = real_t
        self.workers = workers
        self.x_range = x_range
        self.y_range = x_range * (grid_size_y / grid_size_x)
        self.dx = real_t(x_range / grid_size_x)
        self.greens_function_field = self.construct_greens_function_field()
        ghat = sfft.rfftn(self.greens_function_field, workers=workers)
        self.fourier_greens_function_times_dx_squared = ghat * (self.dx ** 2)

    def construct_greens_function_field(self):
        """reference UnboundedPoissonSolverMPI2D.py:60-101."""
        t = self.real_t
        dx = self.dx
        x = np.linspace(0 * dx, (2 * self.nx - 1) * dx, 2 * self.nx).astype(t)
        y = np.linspace(0 * dx, (2 * self.ny - 1) * dx, 2 * self.ny).astype(t)
        yy, xx = np.meshgrid(y, x, indexing="ij")
        r = np.sqrt(
            np.minimum(xx, 2 * self.x_range - xx) ** 2
            + np.minimum(yy, 2 * self.y_range - yy) ** 2
        )
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            g = -np.log(r) / (2 * np.pi)
        g[0, 0] = -(2 * np.log(dx / np.sqrt(np.pi)) - 1) / (4 * np.pi)
        return g.astype(t)

    def solve(self, solution_field, rhs_field, gs):
        ny, nx = self.ny, self.nx
        doubled = np.zeros((2 * ny, 2 * nx), dtype=self.real_t)
        inner = (slice(gs, -gs),) * 2 if gs > 0 else (slice(None),) * 2
        doubled[:ny, :nx] = rhs_field[inner]
        fourier = sfft.rfftn(doubled, workers=self.workers)
        conv = fourier * self.fourier_greens_function_times_dx_squared
        back = sfft.irfftn(conv, s=doubled.shape, workers=self.workers)
        solution_field[inner] = back[:ny, :nx]
