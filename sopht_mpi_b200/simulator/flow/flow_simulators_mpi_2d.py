"""2D unbounded flow simulator with the reference's constructor, attributes and step
ordering (``sopht_mpi/simulator/flow/flow_simulators_mpi_2d.py:21-328``); fields live on
the GPU, every operator is a ``libsophtb200`` kernel (2D instantiations of the same
kernels the 3D simulator uses).  The 512x256 float64 configuration this twin exists for
moves ~1 MiB per field, so a step is bound by launch latency, not by HBM bandwidth.
"""
import numpy as np
import torch

from ...numeric.eulerian_grid_ops import (
    UnboundedPoissonSolverMPI2D,
    gen_add_fixed_val_pyst_kernel_2d,
    gen_advection_timestep_euler_forward_conservative_eno3_pyst_mpi_kernel_2d,
    gen_diffusion_timestep_euler_forward_pyst_mpi_kernel_2d,
    gen_outplane_field_curl_pyst_mpi_kernel_2d,
    gen_penalise_field_boundary_pyst_mpi_kernel_2d,
    gen_set_fixed_val_pyst_kernel_2d,
    gen_update_vorticity_from_velocity_forcing_pyst_mpi_kernel_2d,
)
from ...numeric.eulerian_grid_ops.ops import OpContext
from ...utils import MPI, MPIConstruct2D, MPIGhostCommunicator2D, logger
from ...utils.device import dptr, zeros, zeros_like
from ...utils.precision import get_test_tol


class UnboundedFlowSimulator2D:
    """Class for the GPU 2D unbounded flow simulator"""

    def __init__(
        self,
        grid_size,
        x_range,
        kinematic_viscosity,
        time=0.0,
        CFL=0.1,
        flow_type="passive_scalar",
        with_free_stream_flow=False,
        real_t=np.float32,
        rank_distribution=None,
        ghost_size=2,
        **kwargs,
    ):
        self.grid_dim = 2
        self.grid_size = grid_size
        self.grid_size_y, self.grid_size_x = self.grid_size
        self.x_range = x_range
        self.real_t = real_t
        self.flow_type = flow_type
        self.with_free_stream_flow = with_free_stream_flow
        self.kinematic_viscosity = kinematic_viscosity
        self.CFL = CFL
        self.time = time
        supported_flow_types = ["passive_scalar", "navier_stokes", "navier_stokes_with_forcing"]
        if self.flow_type not in supported_flow_types:
            raise ValueError("Invalid flow type given")
        if self.flow_type == "passive_scalar" and self.with_free_stream_flow:
            raise ValueError("Free stream flow not defined for passive advection diffusion!")
        self.rank_distribution = rank_distribution
        self.ghost_size = ghost_size
        self.poisson_backend = kwargs.get("poisson_backend", "auto")

        self.init_mpi()
        self.init_domain()
        self.init_fields()
        if self.flow_type in ["navier_stokes", "navier_stokes_with_forcing"]:
            self.penalty_zone_width = kwargs.get("penalty_zone_width", 2)
        self.compile_kernels()
        self.finalise_flow_timestep()

    def init_mpi(self):
        self.mpi_construct = MPIConstruct2D(
            grid_size_y=self.grid_size_y,
            grid_size_x=self.grid_size_x,
            real_t=self.real_t,
            rank_distribution=self.rank_distribution,
        )
        self.mpi_ghost_exchange_communicator = MPIGhostCommunicator2D(
            ghost_size=self.ghost_size,
            mpi_construct=self.mpi_construct,
            full_exchange=False,
        )
        self.device = self.mpi_construct.device

    def init_domain(self):
        """Local domain (with ghost cells); reference :110-146."""
        self.y_range = self.x_range * self.grid_size_y / self.grid_size_x
        self.dx = self.real_t(self.x_range / self.grid_size_x)
        eul_grid_shift = self.dx / 2.0
        ghost_grid_shift = self.ghost_size * self.dx
        local_grid_size = self.mpi_construct.local_grid_size
        substart_idx = self.mpi_construct.grid.coords * local_grid_size
        subend_idx = substart_idx + local_grid_size
        substart_y, substart_x = substart_idx * self.dx
        subend_y, subend_x = subend_idx * self.dx
        ny, nx = local_grid_size
        gs = self.ghost_size
        self.local_x = np.linspace(eul_grid_shift + substart_x - ghost_grid_shift,
                                   subend_x - eul_grid_shift + ghost_grid_shift,
                                   nx + 2 * gs).astype(self.real_t)
        self.local_y = np.linspace(eul_grid_shift + substart_y - ghost_grid_shift,
                                   subend_y - eul_grid_shift + ghost_grid_shift,
                                   ny + 2 * gs).astype(self.real_t)
        # flipud so that the position field follows the VectorField convention (index 0 = x)
        self.position_field = np.flipud(np.meshgrid(self.local_y, self.local_x, indexing="ij"))
        self.local_grid_size_with_ghost = local_grid_size + 2 * self.ghost_size
        logger.info(
            "==============================================="
            f"\n{self.grid_dim}D flow domain initialized with:"
            f"\nX axis from 0.0 to {self.x_range}"
            f"\nY axis from 0.0 to {self.y_range}"
            "\nPlease initialize bodies within these bounds!"
            "\n===============================================")

    def init_fields(self):
        shape = tuple(int(s) for s in self.local_grid_size_with_ghost)
        self.primary_scalar_field = zeros(shape, self.real_t, self.device)
        self.velocity_field = zeros((self.grid_dim,) + shape, self.real_t, self.device)
        # one buffer for advection, diffusion and the velocity magnitude
        self.buffer_scalar_field = zeros_like(self.primary_scalar_field)
        if self.flow_type in ["navier_stokes", "navier_stokes_with_forcing"]:
            self.vorticity_field = self.primary_scalar_field.view()
            self.stream_func_field = zeros_like(self.vorticity_field)
        if self.flow_type == "navier_stokes_with_forcing":
            self.eul_grid_forcing_field = zeros_like(self.velocity_field)
        self._reduce_dev = torch.zeros(1, dtype=torch.float64, device=self.device)

    def compile_kernels(self):
        """Bind the operators this flow type needs; reference :165-240."""
        common = dict(real_t=self.real_t, mpi_construct=self.mpi_construct,
                      ghost_exchange_communicator=self.mpi_ghost_exchange_communicator)
        self._ctx = OpContext(self.real_t, self.mpi_construct, self.mpi_ghost_exchange_communicator)
        self.diffusion_timestep = gen_diffusion_timestep_euler_forward_pyst_mpi_kernel_2d(**common)
        self.advection_timestep = (
            gen_advection_timestep_euler_forward_conservative_eno3_pyst_mpi_kernel_2d(**common))
        if self.flow_type in ["navier_stokes", "navier_stokes_with_forcing"]:
            self.unbounded_poisson_solver = UnboundedPoissonSolverMPI2D(
                grid_size_y=self.grid_size_y,
                grid_size_x=self.grid_size_x,
                x_range=self.x_range,
                real_t=self.real_t,
                mpi_construct=self.mpi_construct,
                ghost_size=self.ghost_size,
                backend=self.poisson_backend,
            )
            self.curl = gen_outplane_field_curl_pyst_mpi_kernel_2d(**common)
            self.penalise_field_towards_boundary = gen_penalise_field_boundary_pyst_mpi_kernel_2d(
                width=self.penalty_zone_width,
                dx=self.dx,
                x_grid_field=self.position_field[0],
                y_grid_field=self.position_field[1],
                **common,
            )
        if self.flow_type == "navier_stokes_with_forcing":
            self.update_vorticity_from_velocity_forcing = (
                gen_update_vorticity_from_velocity_forcing_pyst_mpi_kernel_2d(**common))
            self.set_field = gen_set_fixed_val_pyst_kernel_2d(real_t=self.real_t, field_type="vector")
        if self.with_free_stream_flow:
            add_fixed_val = gen_add_fixed_val_pyst_kernel_2d(real_t=self.real_t, field_type="vector")

            def update_velocity_with_free_stream(free_stream_velocity):
                add_fixed_val(sum_field=self.velocity_field, vector_field=self.velocity_field,
                              fixed_vals=free_stream_velocity)
        else:
            def update_velocity_with_free_stream(free_stream_velocity):
                ...

        self.update_velocity_with_free_stream = update_velocity_with_free_stream

    def finalise_flow_timestep(self):
        self.flow_time_step = self.advection_and_diffusion_timestep
        if self.flow_type == "navier_stokes":
            self.flow_time_step = self.navier_stokes_timestep
        elif self.flow_type == "navier_stokes_with_forcing":
            self.flow_time_step = self.navier_stokes_with_forcing_timestep

    def update_simulator_time(self, dt):
        self.time += dt

    def time_step(self, dt, **kwargs):
        self.flow_time_step(dt=dt, **kwargs)
        self.update_simulator_time(dt=dt)

    def advection_and_diffusion_timestep(self, dt, **kwargs):
        """reference :255-266"""
        self.advection_timestep(
            field=self.primary_scalar_field,
            advection_flux=self.buffer_scalar_field,
            velocity=self.velocity_field,
            dt_by_dx=self.real_t(dt / self.dx),
        )
        self.diffusion_timestep(
            field=self.primary_scalar_field,
            diffusion_flux=self.buffer_scalar_field,
            nu_dt_by_dx2=self.real_t(self.kinematic_viscosity * dt / self.dx / self.dx),
        )

    def compute_velocity_from_vorticity(self):
        """penalise -> Poisson -> u = curl(psi z)/(2dx); reference :268-277"""
        self.penalise_field_towards_boundary(field=self.vorticity_field)
        self.unbounded_poisson_solver.solve(solution_field=self.stream_func_field,
                                            rhs_field=self.vorticity_field)
        self.curl(curl=self.velocity_field, field=self.stream_func_field,
                  prefactor=self.real_t(0.5 / self.dx))

    def navier_stokes_timestep(self, dt, free_stream_velocity=(0.0, 0.0)):
        """reference :279-282"""
        self.advection_and_diffusion_timestep(dt=dt)
        self.compute_velocity_from_vorticity()
        self.update_velocity_with_free_stream(free_stream_velocity=free_stream_velocity)

    def navier_stokes_with_forcing_timestep(self, dt, free_stream_velocity=(0.0, 0.0)):
        """reference :284-293"""
        self.update_vorticity_from_velocity_forcing(
            vorticity_field=self.vorticity_field,
            velocity_forcing_field=self.eul_grid_forcing_field,
            prefactor=self.real_t(dt / (2 * self.dx)),
        )
        self.navier_stokes_timestep(dt=dt, free_stream_velocity=free_stream_velocity)
        self.set_field(vector_field=self.eul_grid_forcing_field, fixed_vals=[0.0] * self.grid_dim)

    def _reduce(self, name, field, ncomp):
        ctx = self._ctx
        ctx.call(name, ctx.gref, dptr(field.tensor), ncomp, dptr(self._reduce_dev), ctx.stream())
        return float(self._reduce_dev.item())

    def compute_stable_timestep(self, dt_prefac=1, precision="single"):
        """reference :295-318: max over the interior of |u_x| + |u_y| (one reduction kernel)."""
        max_vel = self.real_t(self._reduce("sb200_max_abs_sum", self.velocity_field, self.grid_dim))
        dt = min(
            self.CFL * self.dx / (max_vel + get_test_tol(precision)),
            0.9 * self.dx ** 2 / (2 * self.grid_dim) / (self.kinematic_viscosity),
        )
        dt = self.mpi_construct.grid.allreduce(dt, op=MPI.MIN)
        return dt * dt_prefac

    def get_max_vorticity(self):
        """reference :320-328"""
        local = self._reduce("sb200_max", self.vorticity_field, 1)
        return self.mpi_construct.grid.allreduce(local, op=MPI.MAX)
