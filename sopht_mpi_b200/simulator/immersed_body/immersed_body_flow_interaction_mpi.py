"""Immersed body <-> flow interactors (reference
``sopht_mpi/simulator/immersed_body/immersed_body_flow_interaction_mpi.py:10-202``,
``cosserat_rod/cosserat_rod_flow_interaction_mpi.py:10-80``,
``rigid_body/rigid_body_flow_interaction_mpi.py:10-80``)."""
import numpy as np

from ...numeric.immersed_boundary_ops import VirtualBoundaryForcingMPI
from ...utils import logger
from ...utils.comm import MPI
from .immersed_body_forcing_grid import EmptyForcingGrid


def _report_lagrangian_resolution(grid_name, max_lag_grid_dx, dx):
    """One log record telling the user how the forcing-point spacing compares with the flow grid
    (the coupling works best for spacings between dx / 2 and 2 dx)."""
    ratio = max_lag_grid_dx / dx
    if ratio > 2.0:
        verdict = ("too coarse for the flow grid (more than 2 dx between forcing points): the body leaks, "
                   "refine the Lagrangian grid")
    elif ratio < 0.5:
        verdict = ("finer than the flow grid needs (less than dx / 2 between forcing points): redundant "
                   "points, coarsen the Lagrangian grid")
    else:
        verdict = "matched to the flow grid"
    logger.warning(f"{grid_name}: max Lagrangian spacing {max_lag_grid_dx:.6g} = {ratio:.3g} dx "
                   f"(dx = {dx:.6g}) is {verdict}")


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
        self.auto_ghosting = bool(auto_ghosting)
        # views: the interactor follows whatever the simulator does to its fields
        self.eul_grid_forcing_field = eul_grid_forcing_field.view()
        self.eul_grid_velocity_field = eul_grid_velocity_field.view()
        self.eul_grid_velocity_field.flags.writeable = False

        spacing = mpi_construct.grid.bcast(self.forcing_grid.get_maximum_lagrangian_grid_spacing(),
                                           root=self.master_rank)
        _report_lagrangian_resolution(type(self.forcing_grid).__name__, spacing, dx)
        area = spacing ** (grid_dim - 1)
        VirtualBoundaryForcingMPI.__init__(
            self, mpi_construct=mpi_construct, ghost_size=mpi_ghost_exchange_communicator.ghost_size,
            virtual_boundary_stiffness_coeff=virtual_boundary_stiffness_coeff * area,
            virtual_boundary_damping_coeff=virtual_boundary_damping_coeff * area, grid_dim=grid_dim, dx=dx,
            eul_grid_coord_shift=eul_grid_coord_shift, interp_kernel_width=interp_kernel_width,
            enable_eul_grid_forcing_reset=enable_eul_grid_forcing_reset, start_time=start_time,
            master_rank=self.master_rank, global_lag_grid_position_field=self.forcing_grid.position_field,
            assume_data_locality=assume_data_locality)
        if not self.auto_ghosting:
            logger.warning("interactor created with auto_ghosting=False: exchange the velocity ghost cells "
                           "yourself before every interaction")

    # ------------------------------------------------------------------ the two interactions
    def _prepare(self):
        """Fresh velocity ghost cells (points near a slab face interpolate across it) and current
        Lagrangian kinematics."""
        if self.auto_ghosting:
            u = self.eul_grid_velocity_field
            u.flags.writeable = True
            self.mpi_ghost_exchange_communicator.exchange_vector_field_init(u)
            self.mpi_ghost_exchange_communicator.exchange_finalise()
            u.flags.writeable = False
        self.forcing_grid.compute_lag_grid_position_field()
        self.forcing_grid.compute_lag_grid_velocity_field()

    def compute_interaction_on_lag_grid(self):
        """Forces on the Lagrangian points only (the body's sub-steps between two flow steps)."""
        self._prepare()
        self.compute_interaction_force_on_lag_grid(
            local_eul_grid_velocity_field=self.eul_grid_velocity_field,
            global_lag_grid_position_field=self.forcing_grid.position_field,
            global_lag_grid_velocity_field=self.forcing_grid.velocity_field)

    def compute_full_interaction(self):
        """Forces on the Lagrangian points AND their reaction spread onto the Eulerian forcing field."""
        self._prepare()
        self.compute_interaction_forcing(
            local_eul_grid_forcing_field=self.eul_grid_forcing_field,
            local_eul_grid_velocity_field=self.eul_grid_velocity_field,
            global_lag_grid_position_field=self.forcing_grid.position_field,
            global_lag_grid_velocity_field=self.forcing_grid.velocity_field)

    __call__ = compute_full_interaction

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
