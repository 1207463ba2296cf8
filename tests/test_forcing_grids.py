"""Closed-form checks of the in-repo forcing grids (the reference imports them from the
un-vendored ``sopht`` package, so there is nothing to compare against but geometry and mechanics):
points lie on the body surface, the advertised maximum spacing bounds the real one, point velocities
are rigid-body velocities, and the force / torque handed to the body balances the point forces."""
import types

import numpy as np
import pytest

from sopht_mpi_b200.simulator import (CircularCylinderForcingGrid, CosseratRodElementCentricForcingGrid,
                                      CosseratRodSurfaceForcingGrid, FlowForces, SphereForcingGrid)


def _rotation(axis, angle):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    k = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * k + (1 - np.cos(angle)) * (k @ k)


def _rigid_body(radius, com, rot=np.eye(3), vel=(0, 0, 0), omega_body=(0, 0, 0)):
    b = types.SimpleNamespace()
    b.radius = radius
    b.position_collection = np.array(com, float).reshape(3, 1)
    b.velocity_collection = np.array(vel, float).reshape(3, 1)
    b.omega_collection = np.array(omega_body, float).reshape(3, 1)
    b.director_collection = np.asarray(rot, float).reshape(3, 3, 1)  # rows: body axes in the lab frame
    return b


def _straight_rod(n_elems, start, direction, normal, length, radius):
    direction, normal = np.asarray(direction, float), np.asarray(normal, float)
    r = types.SimpleNamespace()
    r.n_elems = n_elems
    s = np.linspace(0.0, length, n_elems + 1)
    r.position_collection = np.asarray(start, float).reshape(3, 1) + direction.reshape(3, 1) * s
    r.velocity_collection = np.zeros((3, n_elems + 1))
    r.omega_collection = np.zeros((3, n_elems))
    binormal = np.cross(direction, normal)
    q = np.stack([normal, binormal, direction])  # rows d1, d2, d3
    r.director_collection = np.repeat(q[:, :, None], n_elems, axis=2)
    r.radius = np.full(n_elems, radius) if np.isscalar(radius) else np.asarray(radius, float)
    r.lengths = np.full(n_elems, length / n_elems)
    return r


def test_circular_cylinder_grid_geometry_and_mechanics():
    rot = _rotation([0, 0, 1], 0.4)
    body = _rigid_body(0.03, (0.2, 0.25, 0.0), rot, vel=(0.5, -0.2, 0.0), omega_body=(0, 0, 3.0))
    grid = CircularCylinderForcingGrid(grid_dim=2, rigid_body=body, num_forcing_points=60)
    assert grid.num_lag_nodes == 60 and grid.position_field.shape == (2, 60)
    rel = grid.position_field - body.position_collection[:2]
    assert np.allclose(np.linalg.norm(rel, axis=0), 0.03)
    spacing = np.linalg.norm(np.diff(np.c_[grid.position_field, grid.position_field[:, :1]], axis=1), axis=0)
    assert spacing.max() <= grid.get_maximum_lagrangian_grid_spacing() * (1 + 1e-12)
    assert grid.get_maximum_lagrangian_grid_spacing() == pytest.approx(2 * np.pi * 0.03 / 60)
    # rigid-body velocity: v + omega x r (omega along z, unchanged by an in-plane rotation)
    want = body.velocity_collection[:2] + 3.0 * np.stack([-rel[1], rel[0]])
    assert np.allclose(grid.velocity_field, want)
    # force / torque balance
    rng = np.random.default_rng(0)
    f = rng.normal(size=(2, 60))
    forces, torques = np.ones((3, 1)), np.ones((3, 1))
    grid.transfer_forcing_from_grid_to_body(forces, torques, f)
    assert np.allclose(forces[:, 0], [-f[0].sum(), -f[1].sum(), 0.0])
    assert np.allclose(torques[:, 0], [0, 0, -np.sum(rel[0] * f[1] - rel[1] * f[0])])
    with pytest.raises(ValueError):
        CircularCylinderForcingGrid(grid_dim=3, rigid_body=body, num_forcing_points=8)
    # a body at rest keeps its kinematics version, a moved one does not
    v0 = grid.kinematics_version
    grid.compute_lag_grid_position_field()
    grid.compute_lag_grid_velocity_field()
    assert grid.kinematics_version == v0
    body.position_collection[0, 0] += 0.01
    grid.compute_lag_grid_position_field()
    assert grid.kinematics_version == v0 + 1


def test_sphere_grid_geometry_and_mechanics():
    rot = _rotation([1, 2, 3], 0.7)
    omega_body = np.array([0.3, -1.0, 2.0])
    body = _rigid_body(0.1, (0.25, 0.25, 0.25), rot, vel=(1.0, 0.0, -0.5), omega_body=omega_body)
    n_eq = 96
    grid = SphereForcingGrid(grid_dim=3, rigid_body=body, num_forcing_points_along_equator=n_eq)
    rel = grid.position_field - body.position_collection
    assert np.allclose(np.linalg.norm(rel, axis=0), 0.1)
    # about 4 pi r^2 / h^2 points for a spacing h = 2 pi r / n_eq
    h = 2 * np.pi * 0.1 / n_eq
    assert 0.7 * 4 * np.pi * 0.01 / h ** 2 < grid.num_lag_nodes < 1.3 * 4 * np.pi * 0.01 / h ** 2
    # every point has a neighbour closer than the advertised maximum spacing
    d = np.linalg.norm(rel[:, :, None] - rel[:, None, :], axis=0)
    np.fill_diagonal(d, np.inf)
    assert d.min(axis=0).max() <= grid.get_maximum_lagrangian_grid_spacing() * (1 + 1e-9)
    omega_lab = rot.T @ omega_body
    want = body.velocity_collection + np.cross(omega_lab, rel, axisa=0, axisb=0, axisc=0)
    assert np.allclose(grid.velocity_field, want)
    rng = np.random.default_rng(1)
    f = rng.normal(size=rel.shape)
    forces, torques = np.zeros((3, 1)), np.zeros((3, 1))
    grid.transfer_forcing_from_grid_to_body(forces, torques, f)
    assert np.allclose(forces[:, 0], -f.sum(axis=1))
    torque_lab = -np.cross(rel, f, axisa=0, axisb=0, axisc=0).sum(axis=1)
    assert np.allclose(torques[:, 0], rot @ torque_lab)  # body frame
    with pytest.raises(ValueError):
        SphereForcingGrid(grid_dim=2, rigid_body=body, num_forcing_points_along_equator=8)


@pytest.mark.parametrize("grid_dim", [2, 3])
def test_rod_element_centric_grid(grid_dim):
    rod = _straight_rod(8, (0.1, 0.2, 0.0), (1, 0, 0), (0, 1, 0), 0.8, 0.01)
    rod.velocity_collection[...] = np.linspace(1, 9, 9)
    grid = CosseratRodElementCentricForcingGrid(grid_dim=grid_dim, cosserat_rod=rod)
    assert grid.num_lag_nodes == 8
    x = rod.position_collection[:grid_dim]
    assert np.allclose(grid.position_field, 0.5 * (x[:, 1:] + x[:, :-1]))
    assert np.allclose(grid.velocity_field[0], np.arange(1.5, 9, 1.0))
    assert grid.get_maximum_lagrangian_grid_spacing() == pytest.approx(0.1)
    f = np.arange(grid_dim * 8, dtype=float).reshape(grid_dim, 8)
    forces, torques = np.ones((3, 9)), np.ones((3, 8))
    grid.transfer_forcing_from_grid_to_body(forces, torques, f)
    assert np.allclose(forces[:grid_dim].sum(axis=1), -f.sum(axis=1)) and not torques.any()
    assert np.allclose(forces[0, 0], -0.5 * f[0, 0]) and np.allclose(forces[0, 3], -0.5 * (f[0, 2] + f[0, 3]))


@pytest.mark.parametrize("with_cap", [False, True])
def test_rod_surface_grid(with_cap):
    radius = np.array([0.02, 0.02, 0.01, 0.01, 0.02])
    rod = _straight_rod(5, (0.1, 0.3, 0.3), (0, 0, 1), (1, 0, 0), 0.5, radius)
    rod.velocity_collection[...] = np.array([[0.1], [0.0], [0.2]])
    rod.omega_collection[2] = 1.5  # spin about the rod axis, element frame
    grid = CosseratRodSurfaceForcingGrid(grid_dim=3, cosserat_rod=rod, surface_grid_density_for_largest_element=16,
                                         with_cap=with_cap)
    lateral = 16 * 3 + 8 * 2
    assert grid.num_lag_nodes >= lateral and (grid.num_lag_nodes > lateral) == with_cap
    centre = 0.5 * (rod.position_collection[:, 1:] + rod.position_collection[:, :-1])
    rel = grid.position_field[:, :lateral] - centre[:, grid._elem[:lateral]]
    assert np.allclose(rel[2], 0.0)  # rings lie in the cross-section
    assert np.allclose(np.linalg.norm(rel[:2], axis=0), radius[grid._elem[:lateral]])
    if with_cap:  # cap points sit on the end faces, inside the cross-section
        cap = grid.position_field[:, lateral:]
        assert set(np.round(cap[2], 12)) == {0.3, 0.8}
        assert np.all(np.linalg.norm(cap[:2] - np.array([[0.1], [0.3]]), axis=0) <= 0.02 + 1e-12)
    assert grid.get_maximum_lagrangian_grid_spacing() == pytest.approx(max(0.1, 2 * np.pi * 0.02 / 16))
    arm = grid.moment_arm
    want = np.array([[0.1], [0.0], [0.2]]) + np.cross([0, 0, 1.5], arm, axisa=0, axisb=0, axisc=0)
    assert np.allclose(grid.velocity_field, want)
    rng = np.random.default_rng(2)
    f = rng.normal(size=grid.position_field.shape)
    forces, torques = np.zeros((3, 6)), np.zeros((3, 5))
    grid.transfer_forcing_from_grid_to_body(forces, torques, f)
    assert np.allclose(forces.sum(axis=1), -f.sum(axis=1))
    q = rod.director_collection[..., 0]
    torque_lab = -np.cross(arm, f, axisa=0, axisb=0, axisc=0)
    assert np.allclose(torques.sum(axis=1), q @ torque_lab.sum(axis=1))


def test_flow_forces_adds_the_interactor_loads_to_the_rod():
    calls = []
    interactor = types.SimpleNamespace(body_flow_forces=np.full((3, 4), 2.0), body_flow_torques=np.full((3, 3), -1.0),
                                       compute_flow_forces_and_torques=lambda: calls.append(1))
    rod = types.SimpleNamespace(external_forces=np.ones((3, 4)), external_torques=np.ones((3, 3)))
    ff = FlowForces(interactor)
    ff.apply_forces(rod, time=0.0)
    ff.apply_torques(rod, time=0.0)
    assert calls == [1] and np.all(rod.external_forces == 3.0) and np.all(rod.external_torques == 0.0)
