"""2D unbounded flow simulator: the reference's constructor, public attributes, method names and
step ordering (``sopht_mpi/simulator/flow/flow_simulators_mpi_2d.py:21-328``) on top of
``libsophtb200``.  Every operator is the ``dim = 2`` instantiation of a kernel the 3D simulator uses;
the fields live on the GPU.  The configuration this twin exists for (512 x 256, float64) moves about
1 MiB per field, so a step is bound by launch latency, not by HBM bandwidth.
"""
import numpy as np
import torch

from ...numeric import eulerian_grid_ops as ops
from ...utils import MPI, MPIConstruct2D, MPIGhostCommunicator2D
from ...utils.device import zeros, zeros_like
from ...utils.precision import get_test_tol
from .flow_simulator_common import FlowSimulatorCommon

_FLOW_TYPES = ("passive_scalar", "navier_stokes", "navier_stokes_with_forcing")


class UnboundedFlowSimulator2D(FlowSimulatorCommon):
    """GPU twin of the reference's ``UnboundedFlowSimulator2D`` (explicit Euler steps only)."""

    grid_dim = 2

    def __init__(self, grid_size, x_range, kinematic_viscosity, time=0.0, CFL=0.1,
                 flow_type="passive_scalar", with_free_stream_flow=False, real_t=np.float32,
                 rank_distribution=None, ghost_size=2, **kwargs):
        if flow_type not in _FLOW_TYPES:
            raise ValueError("Invalid flow type given")
        if flow_type == "passive_scalar" and with_free_stream_flow:
            raise ValueError("Free stream flow not defined for passive advection diffusion!")
        self.grid_size = grid_size
        self.grid_size_y, self.grid_size_x = grid_size
        self.x_range, self.real_t = x_range, real_t
        self.kinematic_viscosity, self.CFL, self.time = kinematic_viscosity, CFL, time
        self.flow_type, self.with_free_stream_flow = flow_type, with_free_stream_flow
        self.rank_distribution, self.ghost_size = rank_distribution, ghost_size
        self.poisson_backend = kwargs.get("poisson_backend", "auto")
        self._solves_velocity = flow_type != "passive_scalar"
        if self._solves_velocity:
            self.penalty_zone_width = kwargs.get("penalty_zone_width", 2)
        self.init_mpi()
        self.init_domain()
        self.init_fields()
        self.compile_kernels()
        self.finalise_flow_timestep()

    # ------------------------------------------------------------------ set-up
    def init_mpi(self):
        self.mpi_construct = MPIConstruct2D(grid_size_y=self.grid_size_y, grid_size_x=self.grid_size_x,
                                            real_t=self.real_t, rank_distribution=self.rank_distribution)
        # slabs only: face exchange is all the 2D operators need
        self.mpi_ghost_exchange_communicator = MPIGhostCommunicator2D(
            ghost_size=self.ghost_size, mpi_construct=self.mpi_construct, full_exchange=False)
        self.device = self.mpi_construct.device

    def init_domain(self):
        self._init_local_coordinates()
        # (2, y, x) host array in VectorField order: index 0 holds the x coordinates
        yy, xx = np.meshgrid(self.local_y, self.local_x, indexing="ij")
        self.position_field = np.stack([xx, yy])

    def init_fields(self):
        shape = tuple(int(n) for n in self.local_grid_size_with_ghost)
        self.primary_scalar_field = zeros(shape, self.real_t, self.device)
        self.velocity_field = zeros((2,) + shape, self.real_t, self.device)
        self.buffer_scalar_field = zeros_like(self.primary_scalar_field)  # fluxes and |u| share it
        if self._solves_velocity:
            self.vorticity_field = self.primary_scalar_field.view()
            self.stream_func_field = zeros_like(self.vorticity_field)
        if self.flow_type == "navier_stokes_with_forcing":
            self.eul_grid_forcing_field = zeros_like(self.velocity_field)
        self._reduce_dev = torch.zeros(1, dtype=torch.float64, device=self.device)

    def compile_kernels(self):
        """Bind the operators of this flow type (the reference JIT-compiles them here, :165-240)."""
        common = dict(real_t=self.real_t, mpi_construct=self.mpi_construct,
                      ghost_exchange_communicator=self.mpi_ghost_exchange_communicator)
        self._ctx = ops.OpContext(self.real_t, self.mpi_construct, self.mpi_ghost_exchange_communicator)
        self.diffusion_timestep = ops.gen_diffusion_timestep_euler_forward_pyst_mpi_kernel_2d(**common)
        self.advection_timestep = (
            ops.gen_advection_timestep_euler_forward_conservative_eno3_pyst_mpi_kernel_2d(**common))
        if self._solves_velocity:
            self.unbounded_poisson_solver = ops.UnboundedPoissonSolverMPI2D(
                grid_size_y=self.grid_size_y, grid_size_x=self.grid_size_x, x_range=self.x_range,
                real_t=self.real_t, mpi_construct=self.mpi_construct, ghost_size=self.ghost_size,
                backend=self.poisson_backend)
            self.curl = ops.gen_outplane_field_curl_pyst_mpi_kernel_2d(**common)
            self.penalise_field_towards_boundary = ops.gen_penalise_field_boundary_pyst_mpi_kernel_2d(
                width=self.penalty_zone_width, dx=self.dx, x_grid_field=self.position_field[0],
                y_grid_field=self.position_field[1], **common)
        if self.flow_type == "navier_stokes_with_forcing":
            self.update_vorticity_from_velocity_forcing = (
                ops.gen_update_vorticity_from_velocity_forcing_pyst_mpi_kernel_2d(**common))
            self.set_field = ops.gen_set_fixed_val_pyst_kernel_2d(real_t=self.real_t, field_type="vector")
        add_free_stream = (ops.gen_add_fixed_val_pyst_kernel_2d(real_t=self.real_t, field_type="vector")
                           if self.with_free_stream_flow else None)

        def update_velocity_with_free_stream(free_stream_velocity):
            if add_free_stream is not None:
                add_free_stream(sum_field=self.velocity_field, vector_field=self.velocity_field,
                                fixed_vals=free_stream_velocity)

        self.update_velocity_with_free_stream = update_velocity_with_free_stream

    def finalise_flow_timestep(self):
        self.flow_time_step = {
            "passive_scalar": self.advection_and_diffusion_timestep,
            "navier_stokes": self.navier_stokes_timestep,
            "navier_stokes_with_forcing": self.navier_stokes_with_forcing_timestep,
        }[self.flow_type]

    # ------------------------------------------------------------------ steps (reference :255-293)
    def advection_and_diffusion_timestep(self, dt, **kwargs):
        """ENO3 advection, then diffusion, of the primary scalar (the vorticity in the NS flow types)."""
        q, buf = self.primary_scalar_field, self.buffer_scalar_field
        self.advection_timestep(field=q, advection_flux=buf, velocity=self.velocity_field,
                                dt_by_dx=self.real_t(dt / self.dx))
        self.diffusion_timestep(field=q, diffusion_flux=buf,
                                nu_dt_by_dx2=self.real_t(self.kinematic_viscosity * dt / self.dx / self.dx))

    def compute_velocity_from_vorticity(self):
        """penalise -> unbounded Poisson -> u = curl(psi z) / (2 dx)"""
        self.penalise_field_towards_boundary(field=self.vorticity_field)
        self.unbounded_poisson_solver.solve(solution_field=self.stream_func_field,
                                            rhs_field=self.vorticity_field)
        self.curl(curl=self.velocity_field, field=self.stream_func_field, prefactor=self.real_t(0.5 / self.dx))

    def navier_stokes_timestep(self, dt, free_stream_velocity=(0.0, 0.0)):
        self.advection_and_diffusion_timestep(dt=dt)
        self.compute_velocity_from_vorticity()
        self.update_velocity_with_free_stream(free_stream_velocity=free_stream_velocity)

    def navier_stokes_with_forcing_timestep(self, dt, free_stream_velocity=(0.0, 0.0)):
        self.update_vorticity_from_velocity_forcing(
            vorticity_field=self.vorticity_field, velocity_forcing_field=self.eul_grid_forcing_field,
            prefactor=self.real_t(dt / (2 * self.dx)))
        self.navier_stokes_timestep(dt=dt, free_stream_velocity=free_stream_velocity)
        self.set_field(vector_field=self.eul_grid_forcing_field, fixed_vals=[0.0, 0.0])

    # ------------------------------------------------------------------ diagnostics (reference :295-328)
    def compute_stable_timestep(self, dt_prefac=1, precision="single"):
        """min(advective limit CFL dx / max(|u_x| + |u_y|), diffusive limit 0.9 dx^2 / (4 nu)) over all
        ranks; the maximum is one interior reduction on the device."""
        max_vel = self.real_t(self._reduce("sb200_max_abs_sum", self.velocity_field, 2))
        dt = min(self.CFL * self.dx / (max_vel + get_test_tol(precision)),
                 0.9 * self.dx ** 2 / (2 * self.grid_dim) / self.kinematic_viscosity)
        return self.mpi_construct.grid.allreduce(dt, op=MPI.MIN) * dt_prefac

    def get_max_vorticity(self):
        local = self._reduce("sb200_max", self.vorticity_field, 1)
        return self.mpi_construct.grid.allreduce(local, op=MPI.MAX)
