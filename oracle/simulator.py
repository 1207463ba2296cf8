"""CPU oracle (TEST INFRASTRUCTURE ONLY): single-domain restatement of the
reference simulators' per-step operator ORDER
(``sopht_mpi/simulator/flow/flow_simulators_mpi_3d.py:351-449``, and the 2D twin
``flow_simulators_mpi_2d.py:260-330``), composed from ``oracle.stencils``,
``oracle.poisson``.  The ordering is what the reference's integration test pins
(``tests/test_simulator/test_flow/test_flow_simulators_3d.py:266-330``).
"""
import numpy as np

from . import stencils as st
from .poisson import UnboundedPoissonSolverOracle3D


def get_test_tol(precision="single"):
    return float(10 * np.finfo(np.float32 if precision == "single" else np.float64).eps)


class FlowSimulatorOracle3D:
    def __init__(self, grid_size, x_range, kinematic_viscosity, CFL=0.1,
                 flow_type="navier_stokes_with_forcing", real_t=np.float32, ghost_size=2,
                 penalty_zone_width=2, with_free_stream_flow=False, filter_vorticity=False,
                 filter_setting_dict=None, fft_workers=1):
        self.grid_size = tuple(grid_size)
        nz, ny, nx = self.grid_size
        self.real_t = real_t
        self.gs = ghost_size
        self.flow_type = flow_type
        self.kinematic_viscosity = kinematic_viscosity
        self.CFL = CFL
        self.x_range = x_range
        self.dx = real_t(x_range / nx)
        self.penalty_zone_width = penalty_zone_width
        self.with_free_stream_flow = with_free_stream_flow
        self.filter_vorticity = filter_vorticity
        self.filter_setting_dict = filter_setting_dict or {"order": 2, "type": "multiplicative"}
        self.time = 0.0
        gs = ghost_size
        dx = self.dx

        def line(n):
            # reference :139-153 on a single rank
            return np.linspace(dx / 2.0 - gs * dx, n * dx - dx / 2.0 + gs * dx, n + 2 * gs).astype(real_t)

        self.local_x, self.local_y, self.local_z = line(nx), line(ny), line(nz)
        shape = (nz + 2 * gs, ny + 2 * gs, nx + 2 * gs)
        self.velocity_field = np.zeros((3,) + shape, dtype=real_t)
        self.buffer_scalar_field = np.zeros(shape, dtype=real_t)
        if flow_type == "passive_scalar":
            self.primary_scalar_field = np.zeros(shape, dtype=real_t)
        else:
            self.primary_vector_field = np.zeros((3,) + shape, dtype=real_t)
        if flow_type in ("navier_stokes", "navier_stokes_with_forcing"):
            self.vorticity_field = self.primary_vector_field
            self.stream_func_field = np.zeros_like(self.vorticity_field)
            self.buffer_vector_field = np.zeros_like(self.vorticity_field)
            self.poisson = UnboundedPoissonSolverOracle3D(nz, ny, nx, x_range=x_range, real_t=real_t,
                                                          workers=fft_workers)
        if flow_type == "navier_stokes_with_forcing":
            self.eul_grid_forcing_field = np.zeros_like(self.velocity_field)

    # ---- reference :382-393
    def compute_flow_velocity(self, free_stream_velocity):
        gs = self.gs
        st.penalise_field_boundary_mpi(self.vorticity_field, self.penalty_zone_width, self.dx,
                                       self.local_x, self.local_y, self.local_z, gs)
        self.poisson.vector_field_solve(self.stream_func_field, self.vorticity_field, gs)
        st.curl_mpi(self.velocity_field, self.stream_func_field, self.real_t(0.5 / self.dx), gs)
        if self.with_free_stream_flow:
            for c in range(3):
                self.velocity_field[c] += self.real_t(free_stream_velocity[c])

    # ---- reference :395-413
    def rotational_form_navier_stokes_timestep(self, dt, free_stream_velocity):
        gs = self.gs
        st.elementwise_cross_product(self.buffer_vector_field, self.velocity_field,
                                     self.vorticity_field)
        st.update_vorticity_from_velocity_forcing_mpi(
            self.vorticity_field, self.buffer_vector_field, self.real_t(dt / (2 * self.dx)), gs)
        st.diffusion_timestep_mpi(
            self.vorticity_field, self.buffer_scalar_field,
            self.real_t(self.kinematic_viscosity * dt / self.dx / self.dx), gs)
        if self.filter_vorticity:
            st.laplacian_filter_mpi(self.vorticity_field, self.buffer_vector_field[0],
                                    self.buffer_vector_field[1], self.filter_setting_dict["order"],
                                    self.filter_setting_dict["type"], gs)
        self.compute_flow_velocity(free_stream_velocity)

    # ---- reference :415-424
    def navier_stokes_with_forcing_timestep(self, dt, free_stream_velocity):
        st.update_vorticity_from_velocity_forcing_mpi(
            self.vorticity_field, self.eul_grid_forcing_field, self.real_t(dt / (2 * self.dx)),
            self.gs)
        self.rotational_form_navier_stokes_timestep(dt, free_stream_velocity)
        self.eul_grid_forcing_field[...] = 0

    # ---- reference :356-380
    def advection_and_diffusion_timestep(self, dt):
        field = (self.primary_scalar_field if self.flow_type == "passive_scalar"
                 else self.primary_vector_field)
        st.advection_timestep_mpi(field, self.buffer_scalar_field, self.velocity_field,
                                  self.real_t(dt / self.dx), self.gs)
        st.diffusion_timestep_mpi(field, self.buffer_scalar_field,
                                  self.real_t(self.kinematic_viscosity * dt / self.dx / self.dx),
                                  self.gs)

    def time_step(self, dt, free_stream_velocity=(0.0, 0.0, 0.0)):
        if self.flow_type == "navier_stokes_with_forcing":
            self.navier_stokes_with_forcing_timestep(dt, free_stream_velocity)
        elif self.flow_type == "navier_stokes":
            self.rotational_form_navier_stokes_timestep(dt, free_stream_velocity)
        else:
            self.advection_and_diffusion_timestep(dt)
        self.time += dt

    # ---- reference :426-449
    def compute_stable_timestep(self, dt_prefac=1, precision="single"):
        gs = self.gs
        tol = get_test_tol(precision)
        mag = np.sum(np.fabs(self.velocity_field), axis=0)
        dt = min(
            self.CFL * self.dx / (np.amax(mag[gs:-gs, gs:-gs, gs:-gs]) + tol),
            0.9 * self.dx ** 2 / 6 / (self.kinematic_viscosity + tol),
        )
        return dt * dt_prefac


class FlowSimulatorOracle2D:
    """Single-domain restatement of ``UnboundedFlowSimulator2D``'s step order
    (reference ``flow_simulators_mpi_2d.py:255-328``)."""

    def __init__(self, grid_size, x_range, kinematic_viscosity, CFL=0.1,
                 flow_type="navier_stokes_with_forcing", real_t=np.float64, ghost_size=2,
                 penalty_zone_width=2, with_free_stream_flow=False, fft_workers=1):
        from . import stencils_2d as st2
        from .poisson import UnboundedPoissonSolverOracle2D

        self.st2 = st2
        self.grid_size = tuple(grid_size)
        ny, nx = self.grid_size
        self.real_t = real_t
        self.gs = ghost_size
        self.flow_type = flow_type
        self.kinematic_viscosity = kinematic_viscosity
        self.CFL = CFL
        self.x_range = x_range
        self.dx = real_t(x_range / nx)
        self.penalty_zone_width = penalty_zone_width
        self.with_free_stream_flow = with_free_stream_flow
        self.time = 0.0
        gs, dx = ghost_size, self.dx

        def line(n):  # reference :124-133 on a single rank
            return np.linspace(dx / 2.0 - gs * dx, n * dx - dx / 2.0 + gs * dx, n + 2 * gs).astype(real_t)

        self.local_x, self.local_y = line(nx), line(ny)
        shape = (ny + 2 * gs, nx + 2 * gs)
        self.primary_scalar_field = np.zeros(shape, dtype=real_t)
        self.velocity_field = np.zeros((2,) + shape, dtype=real_t)
        self.buffer_scalar_field = np.zeros(shape, dtype=real_t)
        if flow_type in ("navier_stokes", "navier_stokes_with_forcing"):
            self.vorticity_field = self.primary_scalar_field
            self.stream_func_field = np.zeros_like(self.vorticity_field)
            self.poisson = UnboundedPoissonSolverOracle2D(ny, nx, x_range=x_range, real_t=real_t)
        if flow_type == "navier_stokes_with_forcing":
            self.eul_grid_forcing_field = np.zeros_like(self.velocity_field)

    def advection_and_diffusion_timestep(self, dt):  # reference :255-266
        st2 = self.st2
        st2.advection_timestep_mpi(self.primary_scalar_field, self.buffer_scalar_field, self.velocity_field,
                                   self.real_t(dt / self.dx), self.gs)
        st2.diffusion_timestep_mpi(self.primary_scalar_field, self.buffer_scalar_field,
                                   self.real_t(self.kinematic_viscosity * dt / self.dx / self.dx), self.gs)

    def compute_velocity_from_vorticity(self):  # reference :268-277
        st2 = self.st2
        st2.penalise_field_boundary_mpi(self.vorticity_field, self.penalty_zone_width, self.dx,
                                        self.local_x, self.local_y, self.gs)
        self.poisson.solve(self.stream_func_field, self.vorticity_field, self.gs)
        st2.outplane_field_curl_mpi(self.velocity_field, self.stream_func_field,
                                    self.real_t(0.5 / self.dx), self.gs)

    def navier_stokes_timestep(self, dt, free_stream_velocity):  # reference :279-282
        self.advection_and_diffusion_timestep(dt)
        self.compute_velocity_from_vorticity()
        if self.with_free_stream_flow:
            for c in range(2):
                self.velocity_field[c] += self.real_t(free_stream_velocity[c])

    def time_step(self, dt, free_stream_velocity=(0.0, 0.0)):
        if self.flow_type == "navier_stokes_with_forcing":  # reference :284-293
            self.st2.update_vorticity_from_velocity_forcing_mpi(
                self.vorticity_field, self.eul_grid_forcing_field, self.real_t(dt / (2 * self.dx)), self.gs)
            self.navier_stokes_timestep(dt, free_stream_velocity)
            self.eul_grid_forcing_field[...] = 0
        elif self.flow_type == "navier_stokes":
            self.navier_stokes_timestep(dt, free_stream_velocity)
        else:
            self.advection_and_diffusion_timestep(dt)
        self.time += dt

    def compute_stable_timestep(self, dt_prefac=1, precision="single"):  # reference :295-318
        gs = self.gs
        mag = np.sum(np.fabs(self.velocity_field), axis=0)
        dt = min(self.CFL * self.dx / (np.amax(mag[gs:-gs, gs:-gs]) + get_test_tol(precision)),
                 0.9 * self.dx ** 2 / 4 / self.kinematic_viscosity)
        return dt * dt_prefac
