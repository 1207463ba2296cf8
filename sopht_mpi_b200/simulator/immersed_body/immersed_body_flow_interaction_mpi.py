"""Immersed body <-> flow interactors (reference
``sopht_mpi/simulator/immersed_body/immersed_body_flow_interaction_mpi.py:10-202``,
``cosserat_rod/cosserat_rod_flow_interaction_mpi.py:10-80``,
``rigid_body/rigid_body_flow_interaction_mpi.py:10-80``)."""
import numpy as np

from ...numeric.immersed_boundary_ops import VirtualBoundaryForcingMPI
from ...utils import logger
from ...utils.comm import MPI
from .immersed_body_forcing_grid import EmptyForcingGrid


_RULE = "=========================================================="


def _warn_about_lagrangian_resolution(grid_type, max_lag_grid_dx, dx):
    """The three resolution messages of the reference, word for word (its tests match on them,
    ``tests/test_simulator/immersed_body/test_immersed_body_interaction_mpi.py:57-94``;
    reference ``immersed_body_flow_interaction_mpi.py:49-79``): the delta function spans two
    flow cells, so a spacing above 2 dx leaks and one below dx / 2 is redundant."""
    logger.warning(f"{_RULE}\nFor {grid_type}:")
    if max_lag_grid_dx > 2 * dx:
        logger.warning(
            f"Eulerian grid spacing (dx): {dx}"
            f"\nMax Lagrangian grid spacing: {max_lag_grid_dx} > 2 * dx"
            "\nThe Lagrangian grid of the body is too coarse relative to"
            "\nthe Eulerian grid of the flow, which can lead to unexpected"
            "\nconvergence. Please make the Lagrangian grid finer.")
    elif max_lag_grid_dx < 0.5 * dx:
        logger.warning(
            f"{_RULE}\n"
            f"Eulerian grid spacing (dx): {dx}"
            f"\nMax Lagrangian grid spacing: {max_lag_grid_dx} < 0.5 * dx"
            "\nThe Lagrangian grid of the body is too fine relative to"
            "\nthe Eulerian grid of the flow, which corresponds to redundant"
            "\nforcing points. Please make the Lagrangian grid coarser.")
    else:
        logger.warning(
            "Lagrangian grid is resolved almost the same as the Eulerian"
            "\ngrid of the flow.")
    logger.warning(_RULE)


class ImmersedBodyFlowInteractionMPI(VirtualBoundaryForcingMPI):
    """Couples one body to the flow (API of the reference's ``ImmersedBodyFlowInteractionMPI``,
    ``immersed_body_flow_interaction_mpi.py:10-202``).  Subclasses create ``forcing_grid`` (the real
    one on ``master_rank``, an ``EmptyForcingGrid`` elsewhere), ``body_flow_forces`` and
    ``body_flow_torques`` first.  The penalty coefficients are given per unit area (3D) / length
    (2D) and scaled here by ``max_lag_grid_dx ** (grid_dim - 1)``."""

    def __init__(self, mpi_construct, mpi_ghost_exchange_communicator, eul_grid_forcing_field,
                 eul_grid_velocity_field, virtual_boundary_stiffness_coeff, virtual_boundary_damping_coeff,
                 dx, grid_dim, eul_grid_coord_shift=None, interp_kernel_width=None,
                 enable_eul_grid_forcing_reset=False, start_time=0.0, assume_data_locality=False,
                 auto_ghosting=True):
        self.mpi_ghost_exchange_communicator = mpi_ghost_exchange_communicator
        # views: the interactor follows whatever the simulator does to its fields
        self.eul_grid_forcing_field = eul_grid_forcing_field.view()
        self.eul_grid_velocity_field = eul_grid_velocity_field.view()
        self.eul_grid_velocity_field.flags.writeable = False

        max_lag_grid_dx = mpi_construct.grid.bcast(
            self.forcing_grid.get_maximum_lagrangian_grid_spacing(), root=self.master_rank)
        _warn_about_lagrangian_resolution(type(self.forcing_grid).__name__, max_lag_grid_dx, dx)
        area = max_lag_grid_dx ** (grid_dim - 1)
        VirtualBoundaryForcingMPI.__init__(
            self, mpi_construct=mpi_construct, ghost_size=mpi_ghost_exchange_communicator.ghost_size,
            virtual_boundary_stiffness_coeff=virtual_boundary_stiffness_coeff * area,
            virtual_boundary_damping_coeff=virtual_boundary_damping_coeff * area, grid_dim=grid_dim, dx=dx,
            eul_grid_coord_shift=eul_grid_coord_shift, interp_kernel_width=interp_kernel_width,
            enable_eul_grid_forcing_reset=enable_eul_grid_forcing_reset, start_time=start_time,
            master_rank=self.master_rank, global_lag_grid_position_field=self.forcing_grid.position_field,
            assume_data_locality=assume_data_locality)
        # the public entry points are bound per instance, as in the reference (:104-121), so user
        # code may swap them
        if auto_ghosting:
            self.compute_full_interaction = self._compute_full_interaction_with_ghosting
            self.compute_interaction_on_lag_grid = self._compute_interaction_on_lag_grid_with_ghosting
        else:
            logger.warning(
                f"{_RULE}\n"
                "Auto ghosting of velocity field is disabled for interactor.\n"
                "Please ensure ghosting is done before calling interactor functions.\n"
                f"{_RULE}")
            self.compute_full_interaction = self._compute_full_interaction_without_ghosting
            self.compute_interaction_on_lag_grid = self._compute_interaction_on_lag_grid_without_ghosting

    def __call__(self):
        self.compute_full_interaction()

    # ------------------------------------------------------------------ the two interactions
    def _ghost_velocity_field_for_interaction(self):
        """Fresh velocity ghost cells: points near a slab face interpolate across it."""
        u = self.eul_grid_velocity_field
        u.flags.writeable = True
        self.mpi_ghost_exchange_communicator.exchange_vector_field_init(u)
        self.mpi_ghost_exchange_communicator.exchange_finalise()
        u.flags.writeable = False

    def _update_lagrangian_kinematics(self):
        self.forcing_grid.compute_lag_grid_position_field()
        self.forcing_grid.compute_lag_grid_velocity_field()
        # optional protocol of this build (see VirtualBoundaryForcingMPI._upload_kinematics)
        self._kinematics_version = getattr(self.forcing_grid, "kinematics_version", None)

    def _compute_interaction_on_lag_grid_without_ghosting(self):
        """Forces on the Lagrangian points only (the body's sub-steps between two flow steps)."""
        self._update_lagrangian_kinematics()
        self.compute_interaction_force_on_lag_grid(
            local_eul_grid_velocity_field=self.eul_grid_velocity_field,
            global_lag_grid_position_field=self.forcing_grid.position_field,
            global_lag_grid_velocity_field=self.forcing_grid.velocity_field)

    def _compute_interaction_on_lag_grid_with_ghosting(self):
        self._ghost_velocity_field_for_interaction()
        self._compute_interaction_on_lag_grid_without_ghosting()

    def _compute_full_interaction_without_ghosting(self):
        """Forces on the Lagrangian points AND their reaction spread onto the Eulerian forcing field."""
        self._update_lagrangian_kinematics()
        self.compute_interaction_forcing(
            local_eul_grid_forcing_field=self.eul_grid_forcing_field,
            local_eul_grid_velocity_field=self.eul_grid_velocity_field,
            global_lag_grid_position_field=self.forcing_grid.position_field,
            global_lag_grid_velocity_field=self.forcing_grid.velocity_field)

    def _compute_full_interaction_with_ghosting(self):
        self._ghost_velocity_field_for_interaction()
        self._compute_full_interaction_without_ghosting()

    def compute_flow_forces_and_torques(self):
        self.compute_interaction_on_lag_grid()
        self.forcing_grid.transfer_forcing_from_grid_to_body(
            body_flow_forces=self.body_flow_forces, body_flow_torques=self.body_flow_torques,
            lag_grid_forcing_field=self.global_lag_grid_forcing_field)

    def get_grid_deviation_error_l2_norm(self, compute_global=True):
        """RMS distance between the forcing points and where the body wants them; the global value is
        assembled on ``master_rank`` and handed to every rank."""
        own = float(np.sum(np.square(self.local_lag_grid_position_mismatch_field)))
        count = self.forcing_grid.num_lag_nodes
        if not compute_global:
            return np.sqrt(own / count)
        grid = self.mpi_construct.grid
        total = grid.reduce(own, op=MPI.SUM, root=self.master_rank)
        rms = np.sqrt(total / count) if self.mpi_construct.rank == self.master_rank else None
        return grid.bcast(rms, root=self.master_rank)


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
