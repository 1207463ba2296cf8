"""3D unbounded flow simulator with the reference's constructor, attributes and step
ordering (``sopht_mpi/simulator/flow/flow_simulators_mpi_3d.py:24-476``); fields
live on the GPU, every operator is a ``libsophtb200`` kernel.
"""
import ctypes
import os
from typing import Callable

import numpy as np
import torch

from ... import _lib
from ...numeric.eulerian_grid_ops import (
    UnboundedPoissonSolverMPI3D,
    gen_add_fixed_val_pyst_kernel_3d,
    gen_advection_timestep_euler_forward_conservative_eno3_pyst_mpi_kernel_3d,
    gen_curl_pyst_mpi_kernel_3d,
    gen_diffusion_timestep_euler_forward_pyst_mpi_kernel_3d,
    gen_divergence_pyst_mpi_kernel_3d,
    gen_elementwise_cross_product_pyst_kernel_3d,
    gen_laplacian_filter_mpi_kernel_3d,
    gen_penalise_field_boundary_pyst_mpi_kernel_3d,
    gen_set_fixed_val_pyst_kernel_3d,
    gen_update_vorticity_from_velocity_forcing_pyst_mpi_kernel_3d,
)
from ...numeric.eulerian_grid_ops.ops import OpContext
from ...utils import MPI, MPIConstruct3D, MPIGhostCommunicator3D, logger
from ...utils.deferred import DeferredScalar
from ...utils.device import current_stream_ptr, dptr, zeros, zeros_like
from ...utils.precision import get_test_tol
from .flow_simulator_common import FlowSimulatorCommon

_EAGER_DT = os.environ.get("SB200_EAGER_DT", "0") == "1"


class UnboundedFlowSimulator3D(FlowSimulatorCommon):
    """Class for the GPU 3D unbounded flow simulator"""

    def __init__(
        self,
        grid_size,
        x_range,
        kinematic_viscosity,
        time=0.0,
        CFL=0.1,
        flow_type="passive_scalar",
        filter_vorticity=False,
        real_t=np.float32,
        rank_distribution=None,
        ghost_size=2,
        **kwargs,
    ):
        self.grid_dim = 3
        self.grid_size = grid_size
        self.grid_size_z, self.grid_size_y, self.grid_size_x = self.grid_size
        self.x_range = x_range
        self.real_t = real_t
        self.flow_type = flow_type
        self.kinematic_viscosity = kinematic_viscosity
        self.CFL = CFL
        self.time = time
        self.filter_vorticity = filter_vorticity
        supported_flow_types = [
            "passive_scalar",
            "passive_vector",
            "navier_stokes",
            "navier_stokes_with_forcing",
        ]
        if self.flow_type not in supported_flow_types:
            raise ValueError("Invalid flow type given")
        self.rank_distribution = rank_distribution
        self.ghost_size = ghost_size
        # B200 build options (not in the reference): fused hot-path kernels, Poisson backend
        self.use_fused_kernels = kwargs.get("use_fused_kernels", True)
        self.poisson_backend = kwargs.get("poisson_backend", "auto")

        self.init_mpi()
        self.init_domain()
        self.init_fields()

        if self.flow_type in ["navier_stokes", "navier_stokes_with_forcing"]:
            self.penalty_zone_width = kwargs.get("penalty_zone_width", 2)
            self.with_free_stream_flow = kwargs.get("with_free_stream_flow", False)
            if self.filter_vorticity:
                # same default as the reference when no filter_setting_dict is passed (:97-111)
                given = kwargs.get("filter_setting_dict")
                self.filter_setting_dict = given if given is not None else {"order": 2, "type": "multiplicative"}
                logger.warning(
                    "vorticity filter on: order {order}, type {type}".format(**self.filter_setting_dict)
                    + ("" if given is not None else " (defaults; pass filter_setting_dict={'order': .., 'type': ..})"))
        self.compile_kernels()
        self.finalise_flow_timestep()

    def init_mpi(self):
        self.mpi_construct = MPIConstruct3D(
            grid_size_z=self.grid_size_z,
            grid_size_y=self.grid_size_y,
            grid_size_x=self.grid_size_x,
            real_t=self.real_t,
            rank_distribution=self.rank_distribution,
        )
        need_full_exchange = self.flow_type == "navier_stokes_with_forcing"
        self.mpi_ghost_exchange_communicator = MPIGhostCommunicator3D(
            ghost_size=self.ghost_size,
            mpi_construct=self.mpi_construct,
            full_exchange=need_full_exchange,
        )
        self.device = self.mpi_construct.device

    def init_domain(self):
        """Local domain (with ghost cells), reference :124-168: coordinate lines instead of the full
        meshgrid (see ``position_field``)."""
        self._init_local_coordinates()
        self._position_field = None

    @property
    def position_field(self):
        """(3, z, y, x) host array, index 0 = x coordinates; built on first use
        (1.6 GB at 512^3, which the hot path never needs)."""
        if self._position_field is None:
            self._position_field = np.flipud(np.array(
                np.meshgrid(self.local_z, self.local_y, self.local_x, indexing="ij")))
        return self._position_field

    def _position_lines_as_fields(self):
        shape = tuple(int(s) for s in self.local_grid_size_with_ghost)
        x = np.broadcast_to(self.local_x[None, None, :], shape)
        y = np.broadcast_to(self.local_y[None, :, None], shape)
        z = np.broadcast_to(self.local_z[:, None, None], shape)
        return x, y, z

    def init_fields(self):
        shape = tuple(int(s) for s in self.local_grid_size_with_ghost)
        self.primary_scalar_field = zeros(shape, self.real_t, self.device)
        self.velocity_field = zeros((self.grid_dim,) + shape, self.real_t, self.device)
        self.buffer_scalar_field = zeros_like(self.primary_scalar_field)
        if self.flow_type in ["passive_vector", "navier_stokes", "navier_stokes_with_forcing"]:
            self.primary_vector_field = zeros_like(self.velocity_field)
            del self.primary_scalar_field
        if self.flow_type in ["navier_stokes", "navier_stokes_with_forcing"]:
            self.vorticity_field = self.primary_vector_field.view()
            self.stream_func_field = zeros_like(self.vorticity_field)
            self.buffer_vector_field = zeros_like(self.vorticity_field)
        if self.flow_type == "navier_stokes_with_forcing":
            self.eul_grid_forcing_field = zeros_like(self.velocity_field)
        self._vorticity_alt = None
        self._forcing_tile_flags = None
        self._one_pass_filter = False
        self._max_abs_vel_dev = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._max_abs_vel_version = None
        # global max |u| of the last fused velocity sweep, reduced over the ranks and copied to pinned
        # host memory right behind the sweep (read by the next compute_stable_timestep)
        self._max_abs_vel_glob = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._max_abs_vel_host = torch.zeros(1, dtype=torch.float64, pin_memory=torch.cuda.is_available())
        self._max_abs_vel_event = torch.cuda.Event() if torch.cuda.is_available() else None
        self._pending_dt = None
        self._reduce_dev = torch.zeros(1, dtype=torch.float64, device=self.device)

    def compile_kernels(self):
        """Bind the operators this flow type needs; reference :198-325."""
        common = dict(mpi_construct=self.mpi_construct,
                      ghost_exchange_communicator=self.mpi_ghost_exchange_communicator,
                      real_t=self.real_t)
        self._ctx = OpContext(self.real_t, self.mpi_construct, self.mpi_ghost_exchange_communicator)
        if self.flow_type == "passive_scalar":
            self.diffusion_timestep = gen_diffusion_timestep_euler_forward_pyst_mpi_kernel_3d(
                field_type="scalar", **common)
            self.advection_timestep = (
                gen_advection_timestep_euler_forward_conservative_eno3_pyst_mpi_kernel_3d(
                    field_type="scalar", **common))
        elif self.flow_type == "passive_vector":
            self.diffusion_timestep = gen_diffusion_timestep_euler_forward_pyst_mpi_kernel_3d(
                field_type="vector", **common)
            self.advection_timestep = (
                gen_advection_timestep_euler_forward_conservative_eno3_pyst_mpi_kernel_3d(
                    field_type="vector", **common))

        if self.flow_type in ["navier_stokes", "navier_stokes_with_forcing"]:
            self.diffusion_timestep = gen_diffusion_timestep_euler_forward_pyst_mpi_kernel_3d(
                field_type="vector", **common)
            self.unbounded_poisson_solver = UnboundedPoissonSolverMPI3D(
                grid_size_z=self.grid_size_z,
                grid_size_y=self.grid_size_y,
                grid_size_x=self.grid_size_x,
                x_range=self.x_range,
                real_t=self.real_t,
                mpi_construct=self.mpi_construct,
                ghost_size=self.ghost_size,
                backend=self.poisson_backend,
            )
            self.curl = gen_curl_pyst_mpi_kernel_3d(**common)
            x_grid, y_grid, z_grid = self._position_lines_as_fields()
            self.penalise_field_towards_boundary = gen_penalise_field_boundary_pyst_mpi_kernel_3d(
                width=self.penalty_zone_width,
                dx=self.dx,
                x_grid_field=x_grid,
                y_grid_field=y_grid,
                z_grid_field=z_grid,
                field_type="vector",
                **common,
            )
            self.elementwise_cross_product = gen_elementwise_cross_product_pyst_kernel_3d(
                real_t=self.real_t)
            self.update_vorticity_from_velocity_forcing = (
                gen_update_vorticity_from_velocity_forcing_pyst_mpi_kernel_3d(**common))
            self.compute_divergence = gen_divergence_pyst_mpi_kernel_3d(**common)

            def filter_vector_field(vector_field):
                ...

            self.filter_vector_field = filter_vector_field
            if self.filter_vorticity and self.filter_setting_dict is not None:
                self._one_pass_filter = (
                    self.use_fused_kernels and self.mpi_construct.size == 1
                    and self.filter_setting_dict["order"] == 1
                    and self.filter_setting_dict["type"] == "multiplicative" and self.ghost_size >= 1)
                self.filter_vector_field = gen_laplacian_filter_mpi_kernel_3d(
                    mpi_construct=self.mpi_construct,
                    ghost_exchange_communicator=self.mpi_ghost_exchange_communicator,
                    filter_order=self.filter_setting_dict["order"],
                    filter_flux_buffer=self.buffer_vector_field[0],
                    field_buffer=self.buffer_vector_field[1],
                    real_t=self.real_t,
                    field_type="vector",
                    filter_type=self.filter_setting_dict["type"],
                )

        if self.flow_type == "navier_stokes_with_forcing":
            self.set_field = gen_set_fixed_val_pyst_kernel_3d(real_t=self.real_t, field_type="vector")
        if self.flow_type in ["navier_stokes", "navier_stokes_with_forcing"]:
            if self.with_free_stream_flow:
                add_fixed_val = gen_add_fixed_val_pyst_kernel_3d(real_t=self.real_t,
                                                                 field_type="vector")

                def update_velocity_with_free_stream(free_stream_velocity):
                    add_fixed_val(sum_field=self.velocity_field, vector_field=self.velocity_field,
                                  fixed_vals=free_stream_velocity)
            else:
                def update_velocity_with_free_stream(free_stream_velocity):
                    ...

            self.update_velocity_with_free_stream = update_velocity_with_free_stream

    def finalise_navier_stokes_timestep(self):
        def default_navier_stokes_timestep(dt, free_stream_velocity):
            ...

        self.navier_stokes_timestep = default_navier_stokes_timestep
        if self.flow_type in ["navier_stokes", "navier_stokes_with_forcing"]:
            self.navier_stokes_timestep = self.rotational_form_navier_stokes_timestep

    def finalise_flow_timestep(self):
        self.finalise_navier_stokes_timestep()
        self.flow_time_step: Callable
        self.flow_time_step = self.scalar_advection_and_diffusion_timestep
        if self.flow_type == "passive_vector":
            self.flow_time_step = self.vector_advection_and_diffusion_timestep
        elif self.flow_type == "navier_stokes":
            self.flow_time_step = self.navier_stokes_timestep
        elif self.flow_type == "navier_stokes_with_forcing":
            self.flow_time_step = self.navier_stokes_with_forcing_timestep

    def scalar_advection_and_diffusion_timestep(self, dt: float, **kwargs) -> None:
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

    def vector_advection_and_diffusion_timestep(self, dt: float, **kwargs) -> None:
        self.advection_timestep(
            vector_field=self.primary_vector_field,
            advection_flux=self.buffer_scalar_field,
            velocity=self.velocity_field,
            dt_by_dx=self.real_t(dt / self.dx),
        )
        self.diffusion_timestep(
            vector_field=self.primary_vector_field,
            diffusion_flux=self.buffer_scalar_field,
            nu_dt_by_dx2=self.real_t(self.kinematic_viscosity * dt / self.dx / self.dx),
        )

    # ------------------------------------------------------------- NS hot path
    def compute_flow_velocity(self, free_stream_velocity, reset_forcing=False):
        """penalise -> Poisson -> u = curl(psi)/(2dx) (+ free stream); reference :382-393.
        With ``use_fused_kernels`` the curl, ring zeroing, free-stream add, forcing reset
        and the max-|u| reduction of the next stable-dt run as one sweep."""
        self.penalise_field_towards_boundary(vector_field=self.vorticity_field)
        self.unbounded_poisson_solver.vector_field_solve(
            solution_vector_field=self.stream_func_field,
            rhs_vector_field=self.vorticity_field,
        )
        if not self.use_fused_kernels:
            self.mpi_ghost_exchange_communicator.mark_stale(self.velocity_field)
            self.curl(curl=self.velocity_field, field=self.stream_func_field,
                      prefactor=self.real_t(0.5 / self.dx))
            self.update_velocity_with_free_stream(free_stream_velocity=free_stream_velocity)
            if reset_forcing:
                self.set_field(vector_field=self.eul_grid_forcing_field,
                               fixed_vals=[0.0] * self.grid_dim)
            return
        ctx = self._ctx
        ctx.exchange_vector(self.stream_func_field.tensor)
        fs = (ctypes.c_double * 3)(0.0, 0.0, 0.0)
        if self.with_free_stream_flow and free_stream_velocity is not None:
            for i in range(3):
                fs[i] = float(free_stream_velocity[i])
        forcing = self.eul_grid_forcing_field.tensor if reset_forcing else None
        ctx.call("sb200_velocity_from_stream_function", ctx.gref, dptr(self.velocity_field.tensor),
                 dptr(self.stream_func_field.tensor), float(self.real_t(0.5 / self.dx)), fs,
                 dptr(forcing), dptr(self._max_abs_vel_dev), ctx.stream())
        self._max_abs_vel_version = self.velocity_field.version
        self.mpi_ghost_exchange_communicator.mark_stale(self.velocity_field)
        self._publish_max_abs_vel()

    def _publish_max_abs_vel(self):
        """(A deferred timestep that nobody has read yet is read first: the pinned scalar is about to be
        overwritten.)  Reduce max |u| over the ranks and start its copy to the host behind the producing kernel, so
        that the next compute_stable_timestep only waits for an event.  dt = min over ranks of a
        decreasing function of the local maximum = that function of the global maximum: one NCCL
        all-reduce (MAX) of the device scalar replaces the reference's host allreduce(MIN) of dt
        (flow_simulators_mpi_3d.py:447-448)."""
        if self._pending_dt is not None:
            self._pending_dt._get()
        self._max_abs_vel_glob.copy_(self._max_abs_vel_dev)
        if self._ctx.distributed:
            import torch.distributed as dist

            dist.all_reduce(self._max_abs_vel_glob, op=dist.ReduceOp.MAX)
        self._max_abs_vel_host.copy_(self._max_abs_vel_glob, non_blocking=True)
        self._max_abs_vel_event.record()

    def rotational_form_navier_stokes_timestep(self, dt, free_stream_velocity=None,
                                               _reset_forcing=False):
        """omega += dt curl(u x omega); diffusion; filter; velocity; reference :395-413."""
        if self.use_fused_kernels and self.ghost_size >= 2:
            self._fused_vorticity_update(dt)
        else:
            velocity_cross_vorticity = self.buffer_vector_field.view()
            self.elementwise_cross_product(
                result_field=velocity_cross_vorticity,
                field_1=self.velocity_field,
                field_2=self.vorticity_field,
            )
            self.update_vorticity_from_velocity_forcing(
                vorticity_field=self.vorticity_field,
                velocity_forcing_field=velocity_cross_vorticity,
                prefactor=self.real_t(dt / (2 * self.dx)),
            )
            self.diffusion_timestep(
                vector_field=self.vorticity_field,
                diffusion_flux=self.buffer_scalar_field,
                nu_dt_by_dx2=self.real_t(self.kinematic_viscosity * dt / self.dx / self.dx),
            )
        if self._one_pass_filter:
            self._filter_vorticity_one_pass()
        else:
            self.filter_vector_field(vector_field=self.vorticity_field)
        self.compute_flow_velocity(free_stream_velocity=free_stream_velocity,
                                   reset_forcing=_reset_forcing)

    def _filter_vorticity_one_pass(self):
        """order-1 multiplicative filter in one out-of-place sweep (csrc/stencils.cu); like the fused
        update, the result lands in the second vorticity allocation and the two swap roles"""
        ctx = self._ctx
        w = self.vorticity_field.tensor
        if self._vorticity_alt is None:
            self._vorticity_alt = torch.empty_like(w)
        alt = self._vorticity_alt
        ctx.call("sb200_laplacian_filter_order1_out_of_place", ctx.gref, dptr(alt), dptr(w), self.grid_dim,
                 dptr(self.buffer_vector_field.tensor[0]), dptr(self.buffer_vector_field.tensor[1]), ctx.stream())
        w.data, alt.data = alt.data, w.data

    def _fused_vorticity_update(self, dt):
        """cross product + curl update + diffusion in ONE streaming kernel (csrc/fused.cu).
        Out of place: the result lands in a second vorticity allocation and the two
        allocations swap roles (every DeviceField wrapping the vorticity tensor follows)."""
        ctx = self._ctx
        w, u = self.vorticity_field.tensor, self.velocity_field.tensor
        if self._vorticity_alt is None:
            self._vorticity_alt = torch.empty_like(w)
        alt = self._vorticity_alt
        p = float(self.real_t(dt / (2 * self.dx)))
        d = float(self.real_t(self.kinematic_viscosity * dt / self.dx / self.dx))
        mz = int(w.shape[1])
        lo, hi = 0, mz
        comm = self.mpi_ghost_exchange_communicator
        if ctx.distributed:
            # ghost planes of omega and u stand in for the reference's exchanges of u x omega and of omega
            # between its three sweeps.  The exchange is started, the output planes that do not read a
            # ghost plane (the sweep reaches two planes up and down) run while it is in flight, the
            # planes next to the slab faces follow once it has landed.
            comm.exchange_vector_field_init(self.vorticity_field)
            # (the interactor has usually just exchanged the velocity: its ghost planes are current)
            if not comm.is_fresh(self.velocity_field):
                comm.exchange_vector_field_init(self.velocity_field)
            reach = 2 + self.ghost_size
            faces = self.mpi_construct.physical_faces
            lo = 0 if faces[0] else reach
            hi = mz if faces[1] else mz - reach
        ctx.call("sb200_vorticity_rhs_fused_3d_range", ctx.gref, dptr(alt), dptr(w), dptr(u), p, d, lo, hi,
                 ctx.stream())
        if ctx.distributed:
            comm.exchange_finalise()
            for z0, z1 in ((0, lo), (hi, mz)):
                if z1 > z0:
                    ctx.call("sb200_vorticity_rhs_fused_3d_range", ctx.gref, dptr(alt), dptr(w), dptr(u), p, d, z0, z1,
                             ctx.stream())
        w.data, alt.data = alt.data, w.data

    def navier_stokes_with_forcing_timestep(self, dt, free_stream_velocity=None):
        """reference :415-424.  Fused path: the forcing field of an immersed body is zero almost
        everywhere, so the update only touches the vorticity where curl(F) != 0 and the closing
        ``F = 0`` only the blocks that held a non-zero value (exact for any F; dense F just saves nothing)."""
        if not (self.use_fused_kernels and self.ghost_size >= 2):
            self.update_vorticity_from_velocity_forcing(
                vorticity_field=self.vorticity_field,
                velocity_forcing_field=self.eul_grid_forcing_field,
                prefactor=self.real_t(dt / (2 * self.dx)),
            )
            self.navier_stokes_timestep(dt=dt, free_stream_velocity=free_stream_velocity,
                                        _reset_forcing=True)
            return
        ctx = self._ctx
        f = self.eul_grid_forcing_field.tensor
        if self._forcing_tile_flags is None:
            self._forcing_tile_flags = torch.zeros(int(ctx.lib.sb200_tile_flag_count(ctx.gref)), dtype=torch.uint8,
                                                   device=self.device)
        ctx.exchange_vector(f)
        ctx.call("sb200_update_vorticity_from_sparse_forcing", ctx.gref, dptr(self.vorticity_field.tensor), dptr(f),
                 float(self.real_t(dt / (2 * self.dx))), dptr(self._forcing_tile_flags), ctx.stream())
        # nothing reads F between here and the end of the step: reset it now
        ctx.call("sb200_clear_flagged_tiles", ctx.gref, dptr(f), self.grid_dim, dptr(self._forcing_tile_flags),
                 ctx.stream())
        self.navier_stokes_timestep(dt=dt, free_stream_velocity=free_stream_velocity)

    # ------------------------------------------------------------ diagnostics
    def compute_stable_timestep(self, dt_prefac=1, precision="single"):
        """reference :426-449; max over interior of sum |u_i| comes from the fused velocity
        sweep when the velocity has not been touched from the host since."""
        tol = get_test_tol(precision)
        if not (self._max_abs_vel_version is not None
                and self._max_abs_vel_version == self.velocity_field.version):
            ctx = self._ctx
            ctx.call("sb200_max_abs_sum", ctx.gref, dptr(self.velocity_field.tensor), self.grid_dim,
                     dptr(self._max_abs_vel_dev), ctx.stream())
            # (not cached: only the fused velocity sweep knows that it was the last writer)
            self._publish_max_abs_vel()

        def resolve():
            self._max_abs_vel_event.synchronize()
            max_vel = self.real_t(float(self._max_abs_vel_host[0]))
            dt = min(
                self.CFL * self.dx / (max_vel + tol),
                0.9 * self.dx ** 2 / (2 * self.grid_dim) / (self.kinematic_viscosity + tol),
            )
            self._pending_dt = None
            return dt * dt_prefac

        if _EAGER_DT:
            return resolve()
        # read when first used (utils/deferred.py): the caller's next kernels, the interaction in the
        # reference's step order, are enqueued before the host waits for this value
        if self._pending_dt is not None:
            self._pending_dt._get()
        self._pending_dt = DeferredScalar(resolve)
        return self._pending_dt

    def get_vorticity_divergence_l2_norm(self):
        """reference :451-469"""
        divergence_field = self.buffer_scalar_field.view()
        self.compute_divergence(divergence=divergence_field, field=self.vorticity_field,
                                inv_dx=(1.0 / self.dx))
        local_sq = self._reduce("sb200_sum_squares", divergence_field, 1)
        total = self.mpi_construct.grid.allreduce(local_sq, op=MPI.SUM)
        return np.sqrt(total) * self.dx ** 1.5

    def get_max_vorticity(self):
        """reference :471-476"""
        local = self._reduce("sb200_max", self.vorticity_field, self.grid_dim)
        return self.mpi_construct.grid.allreduce(local, op=MPI.MAX)
