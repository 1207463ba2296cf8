"""Body-specific forcing grids.

The reference takes these classes from the un-vendored ``sopht`` package
(``sopht.simulator.immersed_body``: ``CircularCylinderForcingGrid``, ``SphereForcingGrid``,
``CosseratRodElementCentricForcingGrid``, ``CosseratRodSurfaceForcingGrid``, ``FlowForces``; used at
``examples/2d_examples/FlowPastCylinderCase/flow_past_cylinder.py:62-76``,
``examples/3d_examples/FlowPastSphereCase/flow_past_sphere_case.py:59-82``,
``examples/3d_examples/FlowPastRodCase/flow_past_rod_case.py:121-138``,
``examples/2d_examples/FlowPastRodCase/flow_past_rod.py:124-141``).  Their source is not in the
reference tree, so what follows is written from the interface the reference relies on
(``immersed_body_forcing_grid.py:1-26``, the interactors) and the geometry the examples describe;
the tests pin the geometry in closed form (points on the body surface, advertised maximum spacing,
rigid-body velocities, force / torque balance), not against ``sopht`` itself (PARITY UNPINNED).

Bodies are duck typed on the pyelastica attributes these grids read:

* rigid bodies: ``position_collection (3, 1)``, ``velocity_collection (3, 1)``,
  ``omega_collection (3, 1)`` (body frame), ``director_collection (3, 3, 1)`` (rows = body axes in
  the lab frame), ``radius``;
* Cosserat rods: ``n_elems``, ``position_collection (3, n + 1)``, ``velocity_collection (3, n + 1)``,
  ``omega_collection (3, n)`` (element frame), ``director_collection (3, 3, n)``, ``radius (n,)``,
  ``lengths (n,)``.

Every grid publishes ``kinematics_version`` (see ``VirtualBoundaryForcingMPI._upload_kinematics``): it
changes whenever the grid's positions or velocities changed, so a body at rest is uploaded once.
"""
import numpy as np

from .immersed_body_forcing_grid import ImmersedBodyForcingGrid


def _fingerprint(*arrays):
    """cheap change detector for the (small) body state arrays"""
    return tuple(np.asarray(a).tobytes() for a in arrays)


class _VersionedGrid(ImmersedBodyForcingGrid):
    def __init__(self, grid_dim, num_lag_nodes):
        super().__init__(grid_dim=grid_dim, num_lag_nodes=num_lag_nodes)
        self.kinematics_version = 0
        self._state_seen = None

    def _body_state(self):
        raise NotImplementedError

    def _refresh_version(self):
        state = _fingerprint(*self._body_state())
        if state != self._state_seen:
            self._state_seen = state
            self.kinematics_version += 1


# ------------------------------------------------------------------------------ rigid bodies
class _RigidBodyForcingGrid(_VersionedGrid):
    """Forcing points fixed in the frame of a rigid body: ``local_frame_relative_position_field``
    ``(grid_dim, N)`` holds them relative to the centre of mass, in body axes."""

    def __init__(self, grid_dim, rigid_body, num_lag_nodes):
        self.rigid_body = rigid_body
        super().__init__(grid_dim=grid_dim, num_lag_nodes=num_lag_nodes)
        self.local_frame_relative_position_field = np.zeros((grid_dim, num_lag_nodes))
        self.global_frame_relative_position_field = np.zeros((grid_dim, num_lag_nodes))

    def _body_state(self):
        b = self.rigid_body
        return b.position_collection, b.velocity_collection, b.omega_collection, b.director_collection

    def _lab_from_body(self):
        d = self.grid_dim
        # rows of the director are the body axes in the lab frame: lab = Q^T body
        return np.asarray(self.rigid_body.director_collection)[:d, :d, 0].T

    def compute_lag_grid_position_field(self):
        d = self.grid_dim
        self.global_frame_relative_position_field[...] = (
            self._lab_from_body() @ self.local_frame_relative_position_field)
        self.position_field[...] = (np.asarray(self.rigid_body.position_collection)[:d, 0:1]
                                    + self.global_frame_relative_position_field)
        self._refresh_version()

    def compute_lag_grid_velocity_field(self):
        """v = v_com + omega x r with omega taken to the lab frame"""
        d = self.grid_dim
        body = self.rigid_body
        omega_lab = np.asarray(body.director_collection)[..., 0].T @ np.asarray(body.omega_collection)[:, 0]
        r = self.global_frame_relative_position_field
        v = np.asarray(body.velocity_collection)[:d, 0:1]
        if d == 2:
            self.velocity_field[0] = v[0] - omega_lab[2] * r[1]
            self.velocity_field[1] = v[1] + omega_lab[2] * r[0]
        else:
            self.velocity_field[...] = v + np.cross(omega_lab, r, axisa=0, axisb=0, axisc=0)
        self._refresh_version()

    def transfer_forcing_from_grid_to_body(self, body_flow_forces, body_flow_torques, lag_grid_forcing_field):
        """The body feels the reaction of what the points exert on the fluid: force = - sum F, torque
        = - sum r x F, the torque expressed in body axes (what pyelastica integrates)."""
        d = self.grid_dim
        f = np.asarray(lag_grid_forcing_field)
        r = self.global_frame_relative_position_field
        body_flow_forces[...] = 0.0
        body_flow_forces[:d, 0] = -np.sum(f, axis=1)
        torque_lab = np.zeros(3)
        if d == 2:
            torque_lab[2] = -np.sum(r[0] * f[1] - r[1] * f[0])
        else:
            torque_lab[:] = -np.sum(np.cross(r, f, axisa=0, axisb=0, axisc=0), axis=1)
        body_flow_torques[...] = 0.0
        body_flow_torques[:, 0] = np.asarray(self.rigid_body.director_collection)[..., 0] @ torque_lab


class CircularCylinderForcingGrid(_RigidBodyForcingGrid):
    """2D: ``num_forcing_points`` points on the circular cross-section of a cylinder whose axis is
    normal to the plane."""

    def __init__(self, grid_dim, rigid_body, num_forcing_points):
        if grid_dim != 2:
            raise ValueError("Invalid grid dimensions. Cylinder forcing grid is only defined for grid_dim=2")
        super().__init__(grid_dim, rigid_body, num_forcing_points)
        dtheta = 2.0 * np.pi / num_forcing_points
        theta = np.linspace(0.5 * dtheta, 2.0 * np.pi - 0.5 * dtheta, num_forcing_points)
        self.local_frame_relative_position_field[0] = rigid_body.radius * np.cos(theta)
        self.local_frame_relative_position_field[1] = rigid_body.radius * np.sin(theta)
        self.compute_lag_grid_position_field()
        self.compute_lag_grid_velocity_field()

    def get_maximum_lagrangian_grid_spacing(self):
        """arc length between neighbouring points"""
        return float(self.rigid_body.radius * 2.0 * np.pi / self.num_lag_nodes)


class SphereForcingGrid(_RigidBodyForcingGrid):
    """3D: points on circles of latitude of a sphere, ``num_forcing_points_along_equator`` on the
    equator and proportionally fewer towards the poles, so that the spacing is about uniform."""

    def __init__(self, grid_dim, rigid_body, num_forcing_points_along_equator):
        if grid_dim != 3:
            raise ValueError("Invalid grid dimensions. Sphere forcing grid is only defined for grid_dim=3")
        n_eq = int(num_forcing_points_along_equator)
        polar = np.linspace(0.0, np.pi, n_eq // 2)
        per_latitude = np.rint(n_eq * np.sin(polar)).astype(int) + 1
        super().__init__(grid_dim, rigid_body, int(per_latitude.sum()))
        self.num_forcing_points_along_equator = n_eq
        radius = float(np.ravel(rigid_body.radius)[0])
        start = 0
        for theta, count in zip(polar, per_latitude):
            phi = np.linspace(0.0, 2.0 * np.pi, count, endpoint=False)
            sl = slice(start, start + count)
            self.local_frame_relative_position_field[0, sl] = radius * np.sin(theta) * np.cos(phi)
            self.local_frame_relative_position_field[1, sl] = radius * np.sin(theta) * np.sin(phi)
            self.local_frame_relative_position_field[2, sl] = radius * np.cos(theta)
            start += count
        self._radius = radius
        self.compute_lag_grid_position_field()
        self.compute_lag_grid_velocity_field()

    def get_maximum_lagrangian_grid_spacing(self):
        """the larger of the spacing along the equator and the spacing between two latitudes"""
        n_eq = self.num_forcing_points_along_equator
        return float(self._radius * max(2.0 * np.pi / n_eq, np.pi / max(n_eq // 2 - 1, 1)))


# ----------------------------------------------------------------------------- Cosserat rods
class CosseratRodElementCentricForcingGrid(_VersionedGrid):
    """One forcing point at the centre of every rod element (slender rods; 2D or 3D)."""

    def __init__(self, grid_dim, cosserat_rod):
        self.cosserat_rod = cosserat_rod
        super().__init__(grid_dim=grid_dim, num_lag_nodes=int(cosserat_rod.n_elems))
        self.compute_lag_grid_position_field()
        self.compute_lag_grid_velocity_field()

    def _body_state(self):
        return self.cosserat_rod.position_collection, self.cosserat_rod.velocity_collection

    def compute_lag_grid_position_field(self):
        x = np.asarray(self.cosserat_rod.position_collection)[:self.grid_dim]
        self.position_field[...] = 0.5 * (x[:, 1:] + x[:, :-1])
        self._refresh_version()

    def compute_lag_grid_velocity_field(self):
        v = np.asarray(self.cosserat_rod.velocity_collection)[:self.grid_dim]
        self.velocity_field[...] = 0.5 * (v[:, 1:] + v[:, :-1])
        self._refresh_version()

    def transfer_forcing_from_grid_to_body(self, body_flow_forces, body_flow_torques, lag_grid_forcing_field):
        """each element force is shared equally by the element's two nodes; no torque (the points sit on
        the centre line)"""
        d = self.grid_dim
        f = np.asarray(lag_grid_forcing_field)
        body_flow_forces[...] = 0.0
        body_flow_forces[:d, 1:] -= 0.5 * f
        body_flow_forces[:d, :-1] -= 0.5 * f
        body_flow_torques[...] = 0.0

    def get_maximum_lagrangian_grid_spacing(self):
        return float(np.amax(self.cosserat_rod.lengths))


class CosseratRodSurfaceForcingGrid(_VersionedGrid):
    """3D: a ring of forcing points on the lateral surface of every element; the thickest element
    carries ``surface_grid_density_for_largest_element`` points, thinner ones proportionally fewer.
    ``with_cap`` adds concentric rings on the two end faces."""

    def __init__(self, grid_dim, cosserat_rod, surface_grid_density_for_largest_element, with_cap=False):
        if grid_dim != 3:
            raise ValueError("Invalid grid dimensions. Cosserat rod surface forcing grid is only defined for "
                             "grid_dim=3")
        self.cosserat_rod = cosserat_rod
        self.with_cap = bool(with_cap)
        radius = np.asarray(cosserat_rod.radius, dtype=float)
        n_elems = int(cosserat_rod.n_elems)
        density = int(surface_grid_density_for_largest_element)
        self.surface_grid_density_for_largest_element = density
        per_elem = np.maximum(np.rint(radius / np.amax(radius) * density).astype(int), 1)
        # (element index, radial fraction, angle) of every point
        elem, frac, angle = [], [], []
        for i in range(n_elems):
            th = np.linspace(0.0, 2.0 * np.pi, per_elem[i], endpoint=False)
            elem.append(np.full(per_elem[i], i))
            frac.append(np.ones(per_elem[i]))
            angle.append(th)
        if self.with_cap:
            for i in (0, n_elems - 1):
                rings = max(per_elem[i] // 6, 1)  # ring spacing about equal to the arc spacing
                for k in range(rings):
                    fr = k / rings
                    m = max(int(np.rint(per_elem[i] * fr)), 1)
                    elem.append(np.full(m, i))
                    frac.append(np.full(m, fr))
                    angle.append(np.linspace(0.0, 2.0 * np.pi, m, endpoint=False))
        self._elem = np.concatenate(elem)
        self._frac = np.concatenate(frac)
        self._angle = np.concatenate(angle)
        # caps sit at the end NODES, lateral rings at the element centres
        self._axial = np.zeros(self._elem.size)  # offset along the element axis in units of half lengths
        if self.with_cap:
            n_lat = int(per_elem.sum())
            caps = np.arange(self._elem.size) >= n_lat
            self._axial[caps & (self._elem == 0)] = -1.0
            self._axial[caps & (self._elem == n_elems - 1)] = 1.0
        super().__init__(grid_dim=grid_dim, num_lag_nodes=int(self._elem.size))
        self.moment_arm = np.zeros((3, self.num_lag_nodes))
        self.compute_lag_grid_position_field()
        self.compute_lag_grid_velocity_field()

    def _body_state(self):
        rod = self.cosserat_rod
        return (rod.position_collection, rod.velocity_collection, rod.omega_collection, rod.director_collection,
                rod.radius)

    def compute_lag_grid_position_field(self):
        rod = self.cosserat_rod
        x = np.asarray(rod.position_collection)
        centre = 0.5 * (x[:, 1:] + x[:, :-1])
        q = np.asarray(rod.director_collection)  # rows: d1, d2, d3 (tangent) of every element
        r = np.asarray(rod.radius, dtype=float)[self._elem] * self._frac
        half = 0.5 * np.asarray(rod.lengths, dtype=float)[self._elem] * self._axial
        self.moment_arm[...] = (r * np.cos(self._angle) * q[0][:, self._elem]
                                + r * np.sin(self._angle) * q[1][:, self._elem]
                                + half * q[2][:, self._elem])
        self.position_field[...] = centre[:, self._elem] + self.moment_arm
        self._refresh_version()

    def compute_lag_grid_velocity_field(self):
        rod = self.cosserat_rod
        v = np.asarray(rod.velocity_collection)
        v_elem = 0.5 * (v[:, 1:] + v[:, :-1])
        q = np.asarray(rod.director_collection)
        omega_lab = np.einsum("jin,jn->in", q, np.asarray(rod.omega_collection))  # Q^T omega, per element
        self.velocity_field[...] = v_elem[:, self._elem] + np.cross(
            omega_lab[:, self._elem], self.moment_arm, axisa=0, axisb=0, axisc=0)
        self._refresh_version()

    def transfer_forcing_from_grid_to_body(self, body_flow_forces, body_flow_torques, lag_grid_forcing_field):
        """element force = - sum of its points' forces, shared by the two nodes; element torque
        = - sum arm x F in the element frame"""
        rod = self.cosserat_rod
        f = np.asarray(lag_grid_forcing_field)
        n_elems = int(rod.n_elems)
        f_elem = np.zeros((3, n_elems))
        t_elem = np.zeros((3, n_elems))
        np.add.at(f_elem, (slice(None), self._elem), f)
        np.add.at(t_elem, (slice(None), self._elem), np.cross(self.moment_arm, f, axisa=0, axisb=0, axisc=0))
        body_flow_forces[...] = 0.0
        body_flow_forces[:, 1:] -= 0.5 * f_elem
        body_flow_forces[:, :-1] -= 0.5 * f_elem
        q = np.asarray(rod.director_collection)
        body_flow_torques[...] = -np.einsum("ijn,jn->in", q, t_elem)

    def get_maximum_lagrangian_grid_spacing(self):
        rod = self.cosserat_rod
        arc = 2.0 * np.pi * float(np.amax(rod.radius)) / self.surface_grid_density_for_largest_element
        return float(max(np.amax(rod.lengths), arc))


class FlowForces:
    """pyelastica forcing (``add_forcing_to(rod).using(FlowForces, interactor)``): every rod sub-step
    asks the interactor for the current flow forces and adds them to the rod's external loads
    (reference usage: ``flow_past_rod_case.py:135-138``)."""

    def __init__(self, body_flow_interactor):
        self.body_flow_interactor = body_flow_interactor

    def apply_forces(self, system, time=0.0):
        self.body_flow_interactor.compute_flow_forces_and_torques()
        system.external_forces += self.body_flow_interactor.body_flow_forces

    def apply_torques(self, system, time=0.0):
        system.external_torques += self.body_flow_interactor.body_flow_torques
