"""Virtual boundary forcing (penalty immersed-boundary coupling).

Mirror of ``sopht_mpi/numeric/immersed_boundary_ops/VirtualBoundaryForcingMPI.py:21-459``:
same constructor, same public buffers (host numpy arrays with the reference's names,
which tests and the restart example read AND assign), same method names.  The
Lagrangian work runs on the device: positions / body velocities / position mismatch
are uploaded, one fused kernel does nearest index + weights + E->L interpolation +
mismatch + penalty force, one kernel spreads the force (L->E), and the small
``(dim, n)`` results are mirrored back to the host buffers.
"""
import numpy as np
import torch

from ...utils.comm import MPI
from ...utils.device import Staged, torch_dtype
from ...utils.mpi_utils_2d import MPILagrangianFieldCommunicator2D
from ...utils.mpi_utils_3d import MPILagrangianFieldCommunicator3D
from ..eulerian_grid_ops.ops import gen_set_fixed_val_pyst_kernel_2d, gen_set_fixed_val_pyst_kernel_3d
from .eulerian_lagrangian_grid_communicator import (
    EulerianLagrangianGridCommunicatorMPI2D,
    EulerianLagrangianGridCommunicatorMPI3D,
)


class VirtualBoundaryForcingMPI:
    def __init__(
        self,
        mpi_construct,
        ghost_size,
        virtual_boundary_stiffness_coeff,
        virtual_boundary_damping_coeff,
        grid_dim,
        dx,
        eul_grid_coord_shift=None,
        interp_kernel_width=None,
        enable_eul_grid_forcing_reset=True,
        start_time=0.0,
        master_rank=0,
        global_lag_grid_position_field=None,
        assume_data_locality=False,
    ):
        if grid_dim != 2 and grid_dim != 3:
            raise ValueError("Invalid grid dimensions for virtual boundary forcing!")
        self.grid_dim = grid_dim
        self.virtual_boundary_stiffness_coeff = virtual_boundary_stiffness_coeff
        self.virtual_boundary_damping_coeff = virtual_boundary_damping_coeff
        self.time = start_time
        self.assume_data_locality = assume_data_locality
        self.eul_grid_real_t = mpi_construct.real_t
        self.lag_grid_real_t = global_lag_grid_position_field.dtype
        if eul_grid_coord_shift is None:
            eul_grid_coord_shift = self.eul_grid_real_t(dx / 2)
        self.interp_kernel_width = interp_kernel_width
        if interp_kernel_width is None:
            self.interp_kernel_width = 2
        self.ghost_size = ghost_size
        if self.interp_kernel_width > ghost_size:
            raise ValueError(
                f"Field ghost size {ghost_size} needs to be larger than "
                f"interpolation kernel width {self.interp_kernel_width}")
        self.mpi_construct = mpi_construct
        self.device = mpi_construct.device
        lag_comm_cls = (MPILagrangianFieldCommunicator2D if grid_dim == 2
                        else MPILagrangianFieldCommunicator3D)
        comm_cls = (EulerianLagrangianGridCommunicatorMPI2D if grid_dim == 2
                    else EulerianLagrangianGridCommunicatorMPI3D)
        if not self.assume_data_locality:
            self.mpi_lagrangian_field_communicator = lag_comm_cls(
                eul_grid_dx=dx,
                eul_grid_coord_shift=eul_grid_coord_shift,
                mpi_construct=self.mpi_construct,
                master_rank=master_rank,
                real_t=self.lag_grid_real_t,
            )
        self.eul_lag_grid_communicator = comm_cls(
            dx=dx,
            eul_grid_coord_shift=eul_grid_coord_shift,
            interp_kernel_width=self.interp_kernel_width,
            real_t=self.eul_grid_real_t,
            n_components=grid_dim,
            mpi_construct=mpi_construct,
            ghost_size=ghost_size,
        )
        if not self.assume_data_locality:
            self.mpi_lagrangian_field_communicator.map_lagrangian_nodes_based_on_position(
                global_lag_positions=global_lag_grid_position_field)
            self.local_num_lag_nodes = self.mpi_lagrangian_field_communicator.local_num_lag_nodes
            self.global_num_lag_nodes = self.mpi_lagrangian_field_communicator.rank_address.shape[-1]
        else:
            self.local_num_lag_nodes = global_lag_grid_position_field.shape[-1]
            self.global_num_lag_nodes = self.local_num_lag_nodes

        self._init_local_buffers(self.local_num_lag_nodes)
        self._init_global_buffers()

        if enable_eul_grid_forcing_reset:
            gen = gen_set_fixed_val_pyst_kernel_2d if grid_dim == 2 else gen_set_fixed_val_pyst_kernel_3d
            self.set_eul_grid_vector_field = gen(real_t=self.eul_grid_real_t, field_type="vector")
            self.compute_interaction_forcing = (
                self.compute_interaction_force_on_eul_and_lag_grid_with_eul_grid_forcing_reset)
        else:
            self.compute_interaction_forcing = self.compute_interaction_force_on_eul_and_lag_grid

    # ------------------------------------------------------------------ buffers
    def _init_global_buffers(self):
        if not self.assume_data_locality:
            self.global_lag_grid_position_mismatch_field = np.zeros(
                (self.grid_dim, self.global_num_lag_nodes), dtype=self.lag_grid_real_t)
            self.global_lag_grid_velocity_mismatch_field = np.zeros_like(
                self.global_lag_grid_position_mismatch_field)
            self.global_lag_grid_forcing_field = np.zeros_like(
                self.global_lag_grid_position_mismatch_field)
        else:
            self.global_lag_grid_position_mismatch_field = (
                self.local_lag_grid_position_mismatch_field.view())
            self.global_lag_grid_velocity_mismatch_field = (
                self.local_lag_grid_velocity_mismatch_field.view())
            self.global_lag_grid_forcing_field = self.local_lag_grid_forcing_field.view()

    def _init_local_buffers(self, num_lag_nodes):
        dim, w = self.grid_dim, self.interp_kernel_width
        self.local_nearest_eul_grid_index_to_lag_grid = np.empty((dim, num_lag_nodes), dtype=int)
        self.local_local_eul_grid_support_of_lag_grid = None  # not materialised (fused on device)
        self.local_interp_weights = np.empty((2 * w,) * dim + (num_lag_nodes,),
                                             dtype=self.lag_grid_real_t)
        self.local_lag_grid_flow_velocity_field = np.zeros((dim, num_lag_nodes),
                                                           dtype=self.lag_grid_real_t)
        self.local_lag_grid_position_mismatch_field = np.zeros_like(
            self.local_lag_grid_flow_velocity_field)
        self.local_lag_grid_velocity_mismatch_field = np.zeros_like(
            self.local_lag_grid_position_mismatch_field)
        self.local_lag_grid_forcing_field = np.zeros_like(self.local_lag_grid_velocity_mismatch_field)
        self.local_lag_grid_position_field = np.zeros_like(self.local_lag_grid_position_mismatch_field)
        self.local_lag_grid_velocity_field = np.zeros_like(self.local_lag_grid_position_field)
        # device staging (one block: pos, vel, dpos | flow_vel, dvel, force)
        lt = torch_dtype(self.lag_grid_real_t)
        n = max(int(num_lag_nodes), 1)
        self._dev_in = torch.zeros((3, dim, n), dtype=lt, device=self.device)
        self._dev_out = torch.zeros((3, dim, n), dtype=lt, device=self.device)
        self._dev_nearest = torch.zeros((dim, n), dtype=torch.int64, device=self.device)
        self._dev_weights = torch.zeros(((2 * w) ** dim, n), dtype=lt, device=self.device)
        pin = torch.cuda.is_available()
        self._host_in = torch.zeros((3, dim, n), dtype=lt, pin_memory=pin)
        self._host_out = torch.zeros((3, dim, n), dtype=lt, pin_memory=pin)
        self._fetch_index_and_weights = False

    def update_buffers(self, global_lag_grid_position_field):
        """reference :238-276"""
        comm = self.mpi_lagrangian_field_communicator
        comm.gather_local_field(global_lag_field=self.global_lag_grid_position_mismatch_field,
                                local_lag_field=self.local_lag_grid_position_mismatch_field)
        comm.gather_local_field(global_lag_field=self.global_lag_grid_velocity_mismatch_field,
                                local_lag_field=self.local_lag_grid_velocity_mismatch_field)
        comm.map_lagrangian_nodes_based_on_position(global_lag_grid_position_field)
        update_buffer_flag = self.local_num_lag_nodes != comm.local_num_lag_nodes
        update_buffer_flag = self.mpi_construct.grid.allreduce(update_buffer_flag, op=MPI.LOR)
        if update_buffer_flag:
            self.local_num_lag_nodes = comm.local_num_lag_nodes
            self._init_local_buffers(self.local_num_lag_nodes)
            comm.scatter_global_field(local_lag_field=self.local_lag_grid_position_mismatch_field,
                                      global_lag_field=self.global_lag_grid_position_mismatch_field)
            comm.scatter_global_field(local_lag_field=self.local_lag_grid_velocity_mismatch_field,
                                      global_lag_field=self.global_lag_grid_velocity_mismatch_field)

    # ---- the three pointwise kernels keep their reference names (host numpy views)
    @staticmethod
    def compute_lag_grid_velocity_mismatch_field(lag_grid_velocity_mismatch_field,
                                                 lag_grid_flow_velocity_field,
                                                 lag_grid_body_velocity_field):
        lag_grid_velocity_mismatch_field[...] = (
            lag_grid_flow_velocity_field - lag_grid_body_velocity_field)

    @staticmethod
    def update_lag_grid_position_mismatch_field_via_euler_forward(
            lag_grid_position_mismatch_field, lag_grid_velocity_mismatch_field, dt):
        lag_grid_position_mismatch_field[...] = (
            lag_grid_position_mismatch_field + dt * lag_grid_velocity_mismatch_field)

    @staticmethod
    def compute_lag_grid_forcing_field(lag_grid_forcing_field, lag_grid_position_mismatch_field,
                                       lag_grid_velocity_mismatch_field,
                                       virtual_boundary_stiffness_coeff,
                                       virtual_boundary_damping_coeff):
        lag_grid_forcing_field[...] = (
            virtual_boundary_stiffness_coeff * lag_grid_position_mismatch_field
            + virtual_boundary_damping_coeff * lag_grid_velocity_mismatch_field)

    # ---------------------------------------------------------------- hot path
    def compute_interaction_force_on_lag_grid(self, local_eul_grid_velocity_field,
                                              global_lag_grid_position_field,
                                              global_lag_grid_velocity_field):
        """reference :333-406"""
        if not self.assume_data_locality:
            self.update_buffers(global_lag_grid_position_field=global_lag_grid_position_field)
            comm = self.mpi_lagrangian_field_communicator
            comm.scatter_global_field(local_lag_field=self.local_lag_grid_position_field,
                                      global_lag_field=global_lag_grid_position_field)
            comm.scatter_global_field(local_lag_field=self.local_lag_grid_velocity_field,
                                      global_lag_field=global_lag_grid_velocity_field)
        else:
            self.local_lag_grid_position_field = global_lag_grid_position_field.view()
            self.local_lag_grid_velocity_field = global_lag_grid_velocity_field.view()

        n = int(self.local_num_lag_nodes)
        if n > 0:
            st = Staged(self.device)
            u = st(local_eul_grid_velocity_field)
            hin = self._host_in.numpy()
            hin[0, :, :n] = self.local_lag_grid_position_field
            hin[1, :, :n] = self.local_lag_grid_velocity_field
            hin[2, :, :n] = self.local_lag_grid_position_mismatch_field
            self._dev_in.copy_(self._host_in, non_blocking=True)
            comm_k = self.eul_lag_grid_communicator
            self._params = comm_k.ib_params(self.lag_grid_real_t,
                                            self.virtual_boundary_stiffness_coeff,
                                            self.virtual_boundary_damping_coeff)
            din, dout = self._dev_in, self._dev_out
            nn = din.shape[-1]
            if nn != n:
                raise RuntimeError("local Lagrangian buffer size mismatch")
            comm_k.interact(self._params, n, u, din[0], din[1], din[2], self._dev_nearest,
                            self._dev_weights if self._fetch_index_and_weights else None,
                            dout[0], dout[1], dout[2])
            self._host_out.copy_(dout, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            hout = self._host_out.numpy()
            self.local_lag_grid_flow_velocity_field[...] = hout[0, :, :n]
            self.local_lag_grid_velocity_mismatch_field[...] = hout[1, :, :n]
            self.local_lag_grid_forcing_field[...] = hout[2, :, :n]
            if self._fetch_index_and_weights:
                self.fetch_index_and_weights()

        if not self.assume_data_locality:
            self.mpi_lagrangian_field_communicator.gather_local_field(
                global_lag_field=self.global_lag_grid_forcing_field,
                local_lag_field=self.local_lag_grid_forcing_field)

    def fetch_index_and_weights(self):
        """Mirror nearest indices / interpolation weights of the last interaction to
        the host buffers (off the hot path; the reference materialises them every call)."""
        n = int(self.local_num_lag_nodes)
        if n == 0:
            return
        self.local_nearest_eul_grid_index_to_lag_grid[...] = self._dev_nearest.cpu().numpy()[:, :n]
        w = self.interp_kernel_width
        self.local_interp_weights[...] = self._dev_weights.cpu().numpy().reshape(
            (2 * w,) * self.grid_dim + (-1,))[..., :n]

    def compute_interaction_force_on_eul_and_lag_grid(self, local_eul_grid_forcing_field,
                                                      local_eul_grid_velocity_field,
                                                      global_lag_grid_position_field,
                                                      global_lag_grid_velocity_field):
        """reference :408-429"""
        self.compute_interaction_force_on_lag_grid(local_eul_grid_velocity_field,
                                                   global_lag_grid_position_field,
                                                   global_lag_grid_velocity_field)
        st = Staged(self.device)
        f = st(local_eul_grid_forcing_field, out=True)
        n = int(self.local_num_lag_nodes)
        comm_k = self.eul_lag_grid_communicator
        if n > 0:
            comm_k.spread(self._params, n, f, self._dev_out[2], self._dev_in[0])
        comm_k.eulerian_grid_ghost_sum(local_field=f)
        st.finish()

    def compute_interaction_force_on_eul_and_lag_grid_with_eul_grid_forcing_reset(
            self, local_eul_grid_forcing_field, local_eul_grid_velocity_field,
            global_lag_grid_position_field, global_lag_grid_velocity_field):
        """reference :431-450"""
        self.set_eul_grid_vector_field(vector_field=local_eul_grid_forcing_field,
                                       fixed_vals=([0] * self.grid_dim))
        self.compute_interaction_force_on_eul_and_lag_grid(
            local_eul_grid_forcing_field, local_eul_grid_velocity_field,
            global_lag_grid_position_field, global_lag_grid_velocity_field)

    def time_step(self, dt):
        """reference :452-459 (O(n) host update of the mirrored mismatch state)"""
        self.update_lag_grid_position_mismatch_field_via_euler_forward(
            lag_grid_position_mismatch_field=self.local_lag_grid_position_mismatch_field,
            lag_grid_velocity_mismatch_field=self.local_lag_grid_velocity_mismatch_field,
            dt=dt)
        self.time += dt
