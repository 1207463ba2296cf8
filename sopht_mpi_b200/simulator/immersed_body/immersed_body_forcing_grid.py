"""Forcing-grid interface the interactors talk to.

The concrete body-specific grids (sphere, cylinder, Cosserat-rod surface ...) live in the
un-vendored ``sopht`` package (``sopht.simulator.immersed_body.*ForcingGrid``) and are out
of scope; any object with this interface plugs in.  ``EmptyForcingGrid`` mirrors
``sopht_mpi/simulator/immersed_body/immersed_body_forcing_grid.py:4-26``.
"""
import numpy as np


class ImmersedBodyForcingGrid:
    """Base: owns ``position_field`` / ``velocity_field`` of shape ``(grid_dim, num_lag_nodes)``."""

    def __init__(self, grid_dim, num_lag_nodes):
        self.grid_dim = grid_dim
        self.num_lag_nodes = num_lag_nodes
        self.position_field = np.zeros((grid_dim, num_lag_nodes))
        self.velocity_field = np.zeros_like(self.position_field)

    def compute_lag_grid_position_field(self):
        raise NotImplementedError

    def compute_lag_grid_velocity_field(self):
        raise NotImplementedError

    def transfer_forcing_from_grid_to_body(self, body_flow_forces, body_flow_torques,
                                           lag_grid_forcing_field):
        raise NotImplementedError

    def get_maximum_lagrangian_grid_spacing(self):
        raise NotImplementedError


class EmptyForcingGrid(ImmersedBodyForcingGrid):
    """Placeholder on ranks that do not own the body."""

    def __init__(self, grid_dim):
        super().__init__(grid_dim=grid_dim, num_lag_nodes=0)

    def compute_lag_grid_position_field(self):
        pass

    def compute_lag_grid_velocity_field(self):
        pass

    def transfer_forcing_from_grid_to_body(self, body_flow_forces, body_flow_torques,
                                           lag_grid_forcing_field):
        pass

    def get_maximum_lagrangian_grid_spacing(self):
        pass


class PrescribedForcingGrid(ImmersedBodyForcingGrid):
    """Lagrangian points with prescribed kinematics (synthetic bodies for benchmarks and
    tests): positions/velocities are whatever the owner wrote into the arrays; forces are
    summed onto a single body node.  With ``static=True`` the grid publishes a
    ``kinematics_version`` that only changes when the owner calls :meth:`mark_moved` after rewriting
    the arrays, so the interactor does not upload an unchanged body again."""

    def __init__(self, grid_dim, position_field, velocity_field=None, max_lag_grid_dx=None,
                 centre=None, static=False):
        super().__init__(grid_dim=grid_dim, num_lag_nodes=position_field.shape[-1])
        self.kinematics_version = 0 if static else None
        self.position_field = np.array(position_field)
        self.velocity_field = (np.zeros_like(self.position_field) if velocity_field is None
                               else np.array(velocity_field, dtype=self.position_field.dtype))
        self.max_lag_grid_dx = max_lag_grid_dx
        self.centre = (self.position_field.mean(axis=1) if centre is None else np.asarray(centre))

    def mark_moved(self):
        if self.kinematics_version is not None:
            self.kinematics_version += 1

    def compute_lag_grid_position_field(self):
        pass

    def compute_lag_grid_velocity_field(self):
        pass

    def transfer_forcing_from_grid_to_body(self, body_flow_forces, body_flow_torques,
                                           lag_grid_forcing_field):
        body_flow_forces[...] = 0.0
        body_flow_forces[: self.grid_dim, 0] = -np.sum(lag_grid_forcing_field, axis=1)
        body_flow_torques[...] = 0.0
        if self.grid_dim == 3:
            arm = self.position_field - self.centre.reshape(3, 1)
            body_flow_torques[:, 0] = np.sum(np.cross(arm, -lag_grid_forcing_field, axis=0), axis=1)

    def get_maximum_lagrangian_grid_spacing(self):
        return self.max_lag_grid_dx
