"""Immersed body <-> flow interactors (reference
``sopht_mpi/simulator/immersed_body/immersed_body_flow_interaction_mpi.py:10-202``,
``cosserat_rod/cosserat_rod_flow_interaction_mpi.py:10-80``,
``rigid_body/rigid_body_flow_interaction_mpi.py:10-80``)."""
import numpy as np

from ...numeric.immersed_boundary_ops import VirtualBoundaryForcingMPI
from ...utils import logger
from ...utils.comm import MPI
from .immersed_body_forcing_grid import EmptyForcingGrid


class ImmersedBodyFlowInteractionMPI(VirtualBoundaryForcingMPI):
    """Base class; derived classes set ``body_flow_forces``, ``body_flow_torques``,
    ``forcing_grid`` and ``master_rank`` before calling this initialiser."""

    def __init__(
        self,
        mpi_construct,
        mpi_ghost_exchange_communicator,
        eul_grid_forcing_field,
        eul_grid_velocity_field,
        virtual_boundary_stiffness_coeff,
        virtual_boundary_damping_coeff,
        dx,
        grid_dim,
        eul_grid_coord_shift=None,
        interp_kernel_width=None,
        enable_eul_grid_forcing_reset=False,
        start_time=0.0,
        assume_data_locality=False,
        auto_ghosting=True,
    ):
        self.mpi_ghost_exchange_communicator = mpi_ghost_exchange_communicator
        self.eul_grid_forcing_field = eul_grid_forcing_field.view()
        self.eul_grid_velocity_field = eul_grid_velocity_field.view()
        self.eul_grid_velocity_field.flags.writeable = False

        max_lag_grid_dx = self.forcing_grid.get_maximum_lagrangian_grid_spacing()
        max_lag_grid_dx = mpi_construct.grid.bcast(max_lag_grid_dx, root=self.master_rank)
        grid_type = type(self.forcing_grid).__name__
        logger.warning(
            "==========================================================\n"
            f"For {grid_type}:")
        if max_lag_grid_dx > 2 * dx:
            logger.warning(
                f"Eulerian grid spacing (dx): {dx}"
                f"\nMax Lagrangian grid spacing: {max_lag_grid_dx} > 2 * dx"
                "\nThe Lagrangian grid of the body is too coarse relative to"
                "\nthe Eulerian grid of the flow, which can lead to unexpected"
                "\nconvergence. Please make the Lagrangian grid finer.")
        elif max_lag_grid_dx < 0.5 * dx:
            logger.warning(
                "==========================================================\n"
                f"Eulerian grid spacing (dx): {dx}"
                f"\nMax Lagrangian grid spacing: {max_lag_grid_dx} < 0.5 * dx"
                "\nThe Lagrangian grid of the body is too fine relative to"
                "\nthe Eulerian grid of the flow, which corresponds to redundant"
                "\nforcing points. Please make the Lagrangian grid coarser.")
        else:
            logger.warning(
                "Lagrangian grid is resolved almost the same as the Eulerian"
                "\ngrid of the flow.")
        logger.warning("==========================================================")

        virtual_boundary_stiffness_coeff *= max_lag_grid_dx ** (grid_dim - 1)
        virtual_boundary_damping_coeff *= max_lag_grid_dx ** (grid_dim - 1)

        super().__init__(
            mpi_construct=mpi_construct,
            ghost_size=self.mpi_ghost_exchange_communicator.ghost_size,
            virtual_boundary_stiffness_coeff=virtual_boundary_stiffness_coeff,
            virtual_boundary_damping_coeff=virtual_boundary_damping_coeff,
            grid_dim=grid_dim,
            dx=dx,
            eul_grid_coord_shift=eul_grid_coord_shift,
            interp_kernel_width=interp_kernel_width,
            enable_eul_grid_forcing_reset=enable_eul_grid_forcing_reset,
            start_time=start_time,
            master_rank=self.master_rank,
            global_lag_grid_position_field=self.forcing_grid.position_field,
            assume_data_locality=assume_data_locality,
        )

        if auto_ghosting:
            self.compute_full_interaction = self._compute_full_interaction_with_ghosting
            self.compute_interaction_on_lag_grid = self._compute_interaction_on_lag_grid_with_ghosting
        else:
            logger.warning(
                "==========================================================\n"
                "Auto ghosting of velocity field is disabled for interactor.\n"
                "Please ensure ghosting is done before calling interactor functions.\n"
                "==========================================================")
            self.compute_full_interaction = self._compute_full_interaction_without_ghosting
            self.compute_interaction_on_lag_grid = self._compute_interaction_on_lag_grid_without_ghosting

    def __call__(self):
        self.compute_full_interaction()

    def _ghost_velocity_field_for_interaction(self):
        self.eul_grid_velocity_field.flags.writeable = True
        self.mpi_ghost_exchange_communicator.exchange_vector_field_init(self.eul_grid_velocity_field)
        self.mpi_ghost_exchange_communicator.exchange_finalise()
        self.eul_grid_velocity_field.flags.writeable = False

    def _compute_interaction_on_lag_grid_without_ghosting(self):
        self.forcing_grid.compute_lag_grid_position_field()
        self.forcing_grid.compute_lag_grid_velocity_field()
        self.compute_interaction_force_on_lag_grid(
            local_eul_grid_velocity_field=self.eul_grid_velocity_field,
            global_lag_grid_position_field=self.forcing_grid.position_field,
            global_lag_grid_velocity_field=self.forcing_grid.velocity_field,
        )

    def _compute_interaction_on_lag_grid_with_ghosting(self):
        self._ghost_velocity_field_for_interaction()
        self._compute_interaction_on_lag_grid_without_ghosting()

    def _compute_full_interaction_without_ghosting(self):
        self.forcing_grid.compute_lag_grid_position_field()
        self.forcing_grid.compute_lag_grid_velocity_field()
        self.compute_interaction_forcing(
            local_eul_grid_forcing_field=self.eul_grid_forcing_field,
            local_eul_grid_velocity_field=self.eul_grid_velocity_field,
            global_lag_grid_position_field=self.forcing_grid.position_field,
            global_lag_grid_velocity_field=self.forcing_grid.velocity_field,
        )

    def _compute_full_interaction_with_ghosting(self):
        self._ghost_velocity_field_for_interaction()
        self._compute_full_interaction_without_ghosting()

    def compute_flow_forces_and_torques(self):
        self.compute_interaction_on_lag_grid()
        self.forcing_grid.transfer_forcing_from_grid_to_body(
            body_flow_forces=self.body_flow_forces,
            body_flow_torques=self.body_flow_torques,
            lag_grid_forcing_field=self.global_lag_grid_forcing_field,
        )

    def get_grid_deviation_error_l2_norm(self, compute_global=True):
        if not compute_global:
            return np.linalg.norm(self.local_lag_grid_position_mismatch_field) / np.sqrt(
                self.forcing_grid.num_lag_nodes)
        local_sq = np.linalg.norm(self.local_lag_grid_position_mismatch_field) ** 2
        total = self.mpi_construct.grid.reduce(local_sq, op=MPI.SUM, root=self.master_rank)
        if self.mpi_construct.rank == self.master_rank:
            total = np.sqrt(total) / np.sqrt(self.forcing_grid.num_lag_nodes)
        return self.mpi_construct.grid.bcast(total, root=self.master_rank)


class _BodyFlowInteraction(ImmersedBodyFlowInteractionMPI):
    def __init__(self, body_kw, body, n_force_cols, n_torque_cols, mpi_construct,
                 mpi_ghost_exchange_communicator, eul_grid_forcing_field, eul_grid_velocity_field,
                 virtual_boundary_stiffness_coeff, virtual_boundary_damping_coeff, dx, grid_dim,
                 forcing_grid_cls, eul_grid_coord_shift, interp_kernel_width,
                 enable_eul_grid_forcing_reset, start_time, master_rank, assume_data_locality,
                 auto_ghosting, forcing_grid_kwargs):
        self.body_flow_forces = np.zeros((3, n_force_cols))
        self.body_flow_torques = np.zeros((3, n_torque_cols))
        self.master_rank = master_rank
        if mpi_construct.rank == self.master_rank:
            self.forcing_grid = forcing_grid_cls(grid_dim=grid_dim, **{body_kw: body},
                                                 **forcing_grid_kwargs)
        else:
            self.forcing_grid = EmptyForcingGrid(grid_dim=grid_dim)
        super().__init__(
            mpi_construct=mpi_construct,
            mpi_ghost_exchange_communicator=mpi_ghost_exchange_communicator,
            eul_grid_forcing_field=eul_grid_forcing_field,
            eul_grid_velocity_field=eul_grid_velocity_field,
            virtual_boundary_stiffness_coeff=virtual_boundary_stiffness_coeff,
            virtual_boundary_damping_coeff=virtual_boundary_damping_coeff,
            dx=dx,
            grid_dim=grid_dim,
            eul_grid_coord_shift=eul_grid_coord_shift,
            interp_kernel_width=interp_kernel_width,
            enable_eul_grid_forcing_reset=enable_eul_grid_forcing_reset,
            start_time=start_time,
            assume_data_locality=assume_data_locality,
            auto_ghosting=auto_ghosting,
        )


class CosseratRodFlowInteraction(_BodyFlowInteraction):
    """reference ``cosserat_rod/cosserat_rod_flow_interaction_mpi.py:10-80``"""

    def __init__(self, mpi_construct, mpi_ghost_exchange_communicator, cosserat_rod,
                 eul_grid_forcing_field, eul_grid_velocity_field, virtual_boundary_stiffness_coeff,
                 virtual_boundary_damping_coeff, dx, grid_dim, forcing_grid_cls,
                 eul_grid_coord_shift=None, interp_kernel_width=None,
                 enable_eul_grid_forcing_reset=False, start_time=0.0, master_rank=0,
                 assume_data_locality=False, auto_ghosting=True, **forcing_grid_kwargs):
        super().__init__("cosserat_rod", cosserat_rod, cosserat_rod.n_elems + 1, cosserat_rod.n_elems,
                         mpi_construct, mpi_ghost_exchange_communicator, eul_grid_forcing_field,
                         eul_grid_velocity_field, virtual_boundary_stiffness_coeff,
                         virtual_boundary_damping_coeff, dx, grid_dim, forcing_grid_cls,
                         eul_grid_coord_shift, interp_kernel_width, enable_eul_grid_forcing_reset,
                         start_time, master_rank, assume_data_locality, auto_ghosting,
                         forcing_grid_kwargs)


class RigidBodyFlowInteractionMPI(_BodyFlowInteraction):
    """reference ``rigid_body/rigid_body_flow_interaction_mpi.py:10-80``"""

    def __init__(self, mpi_construct, mpi_ghost_exchange_communicator, rigid_body,
                 eul_grid_forcing_field, eul_grid_velocity_field, virtual_boundary_stiffness_coeff,
                 virtual_boundary_damping_coeff, dx, grid_dim, forcing_grid_cls,
                 eul_grid_coord_shift=None, interp_kernel_width=None,
                 enable_eul_grid_forcing_reset=False, start_time=0.0, master_rank=0,
                 assume_data_locality=False, auto_ghosting=True, **forcing_grid_kwargs):
        super().__init__("rigid_body", rigid_body, 1, 1, mpi_construct,
                         mpi_ghost_exchange_communicator, eul_grid_forcing_field,
                         eul_grid_velocity_field, virtual_boundary_stiffness_coeff,
                         virtual_boundary_damping_coeff, dx, grid_dim, forcing_grid_cls,
                         eul_grid_coord_shift, interp_kernel_width, enable_eul_grid_forcing_reset,
                         start_time, master_rank, assume_data_locality, auto_ghosting,
                         forcing_grid_kwargs)
