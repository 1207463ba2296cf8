"""Parity of the CUDA path (through the operator API -> C ABI -> sm_100a kernels)
against the CPU oracle on identical seeded inputs.

Tolerances (north star): relative L-inf <= 1e-10 in float64, <= 1e-5 in float32.
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import ib as ib_oracle
from oracle import stencils as st
from oracle.poisson import UnboundedPoissonSolverOracle2D, UnboundedPoissonSolverOracle3D
from oracle.simulator import FlowSimulatorOracle3D

pytestmark = pytest.mark.gpu

TOL = {np.float64: 1e-10, np.float32: 1e-5}
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from sopht_mpi_b200 import _lib

    _lib.load()  # fails loudly if the extension is missing
    return torch.device("cuda", 0)


def _constructs(dim, n, real_t, gs=2):
    from sopht_mpi_b200.utils import (MPIConstruct2D, MPIConstruct3D, MPIGhostCommunicator2D,
                                      MPIGhostCommunicator3D)

    if dim == 3:
        mc = MPIConstruct3D(*n, real_t=real_t)
        gc = MPIGhostCommunicator3D(ghost_size=gs, mpi_construct=mc)
    else:
        mc = MPIConstruct2D(*n, real_t=real_t)
        gc = MPIGhostCommunicator2D(ghost_size=gs, mpi_construct=mc)
    return mc, gc


def _dev(a, device):
    from sopht_mpi_b200.utils.device import DeviceField

    return DeviceField(torch.from_numpy(np.ascontiguousarray(a)).to(device))


@pytest.mark.parametrize("real_t", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("n", [(16, 16, 16), (12, 18, 24)], ids=["cube", "aspect"])
def test_stencil_operators_3d(cuda, real_t, n):
    from sopht_mpi_b200.numeric import eulerian_grid_ops as ops

    gs = 2
    rng = np.random.default_rng(0)
    shape = tuple(v + 2 * gs for v in n)
    mc, gc = _constructs(3, n, real_t)
    kw = dict(real_t=real_t, mpi_construct=mc, ghost_exchange_communicator=gc)
    tol = TOL[real_t]
    w = rng.uniform(size=(3,) + shape).astype(real_t)
    f = rng.uniform(size=(3,) + shape).astype(real_t)
    vel = (rng.uniform(size=(3,) + shape) - 0.5).astype(real_t)

    # update vorticity (device fields AND numpy staging path)
    upd = ops.gen_update_vorticity_from_velocity_forcing_pyst_mpi_kernel_3d(**kw)
    assert upd.kernel_support == 1
    ref = w.copy()
    st.update_vorticity_from_velocity_forcing_mpi(ref, f, 0.37, gs)
    dw = _dev(w, cuda)
    upd(vorticity_field=dw, velocity_forcing_field=_dev(f, cuda), prefactor=real_t(0.37))
    assert _rel(dw, ref) <= tol
    host = w.copy()
    upd(vorticity_field=host, velocity_forcing_field=f, prefactor=real_t(0.37))
    assert _rel(host, ref) <= tol

    curl = ops.gen_curl_pyst_mpi_kernel_3d(**kw)
    ref = rng.uniform(size=(3,) + shape).astype(real_t)
    out = _dev(ref, cuda)
    st.curl_mpi(ref, f, 0.8, gs)
    curl(curl=out, field=_dev(f, cuda), prefactor=real_t(0.8))
    assert _rel(out, ref) <= tol

    diff = ops.gen_diffusion_timestep_euler_forward_pyst_mpi_kernel_3d(field_type="vector", **kw)
    ref, flux = w.copy(), np.zeros(shape, real_t)
    st.diffusion_timestep_mpi(ref, flux, 0.1, gs)
    dw, dfl = _dev(w, cuda), _dev(np.zeros(shape, real_t), cuda)
    diff(vector_field=dw, diffusion_flux=dfl, nu_dt_by_dx2=real_t(0.1))
    assert _rel(dw, ref) <= tol and _rel(dfl, flux) <= 10 * tol

    adv = ops.gen_advection_timestep_euler_forward_conservative_eno3_pyst_mpi_kernel_3d(
        field_type="scalar", **kw)
    assert adv.kernel_support == 2
    ref, flux = w[0].copy(), np.ones(shape, real_t)
    st.advection_timestep_mpi(ref, flux, vel, 0.2, gs)
    dq, dfl = _dev(w[0], cuda), _dev(np.ones(shape, real_t), cuda)
    adv(field=dq, advection_flux=dfl, velocity=_dev(vel, cuda), dt_by_dx=real_t(0.2))
    assert _rel(dq, ref) <= tol

    div = ops.gen_divergence_pyst_mpi_kernel_3d(**kw)
    ref = np.zeros(shape, real_t)
    st.divergence_mpi(ref, w, 3.0, gs)
    dd = _dev(np.zeros(shape, real_t), cuda)
    div(divergence=dd, field=_dev(w, cuda), inv_dx=3.0)
    assert _rel(dd, ref) <= tol

    for ftype in ("multiplicative", "convolution"):
        fb, bb = _dev(np.zeros(shape, real_t), cuda), _dev(np.zeros(shape, real_t), cuda)
        filt = ops.gen_laplacian_filter_mpi_kernel_3d(
            mpi_construct=mc, ghost_exchange_communicator=gc, filter_order=2,
            filter_flux_buffer=fb, field_buffer=bb, real_t=real_t, field_type="vector",
            filter_type=ftype)
        ref = w.copy()
        st.laplacian_filter_mpi(ref, np.zeros(shape, real_t), np.zeros(shape, real_t), 2, ftype, gs)
        dw = _dev(w, cuda)
        filt(vector_field=dw)
        assert _rel(dw, ref) <= tol

    # penalise: bit exact (table of sines computed once in real_t, then products only)
    dx = real_t(1.0 / n[2])

    def line(nl):
        return np.linspace(dx / 2 - gs * dx, nl * dx - dx / 2 + gs * dx, nl + 2 * gs).astype(real_t)

    xg, yg, zg = line(n[2]), line(n[1]), line(n[0])
    pos = np.flipud(np.array(np.meshgrid(zg, yg, xg, indexing="ij")))
    pen = ops.gen_penalise_field_boundary_pyst_mpi_kernel_3d(
        width=2, dx=dx, x_grid_field=pos[0], y_grid_field=pos[1], z_grid_field=pos[2],
        field_type="vector", **kw)
    ref = w.copy()
    st.penalise_field_boundary_mpi(ref, 2, dx, xg, yg, zg, gs)
    dw = _dev(w, cuda)
    pen(vector_field=dw)
    assert np.array_equal(np.asarray(dw), ref)

    with pytest.raises(ValueError):
        ops.gen_diffusion_timestep_euler_forward_pyst_mpi_kernel_3d(field_type="tensor", **kw)


def test_ghost_size_smaller_than_kernel_support_raises(cuda):
    from sopht_mpi_b200.numeric import eulerian_grid_ops as ops

    mc, gc = _constructs(3, (8, 8, 8), np.float32, gs=1)
    with pytest.raises(ValueError):
        ops.gen_advection_flux_conservative_eno3_pyst_mpi_kernel_3d(
            real_t=np.float32, mpi_construct=mc, ghost_exchange_communicator=gc)


@pytest.mark.parametrize("backend", ["cufft", "auto"])
@pytest.mark.parametrize("real_t", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("n", [(16, 16, 16), (8, 16, 32)], ids=["cube", "aspect"])
def test_poisson_3d_against_oracle(cuda, real_t, n, backend):
    from sopht_mpi_b200.numeric.eulerian_grid_ops import UnboundedPoissonSolverMPI3D

    gs = 2
    rng = np.random.default_rng(1)
    shape = tuple(v + 2 * gs for v in n)
    mc, _ = _constructs(3, n, real_t)
    solver = UnboundedPoissonSolverMPI3D(*n, mpi_construct=mc, ghost_size=gs, x_range=1.0,
                                         real_t=real_t, backend=backend)
    oracle = UnboundedPoissonSolverOracle3D(*n, x_range=1.0, real_t=real_t)
    rhs = rng.uniform(size=(3,) + shape).astype(real_t)
    ref = np.zeros_like(rhs)
    oracle.vector_field_solve(ref, rhs, gs)
    out = _dev(np.full_like(rhs, 7.0), cuda)
    solver.vector_field_solve(solution_vector_field=out, rhs_vector_field=_dev(rhs, cuda))
    got = np.asarray(out)
    inner = (slice(None),) + (slice(gs, -gs),) * 3
    # float32: both sides carry ~1e-6 FFT round-off; the bar is 1e-5
    assert _rel(got[inner], ref[inner]) <= (1e-10 if real_t == np.float64 else 1e-5)
    # ghosts of the solution are not touched (reference copies interiors only)
    mask = np.ones(shape, bool)
    mask[gs:-gs, gs:-gs, gs:-gs] = False
    assert np.all(got[0][mask] == 7.0)
    # scalar entry point
    one = _dev(np.zeros(shape, real_t), cuda)
    solver.solve(solution_field=one, rhs_field=_dev(rhs[1], cuda))
    assert _rel(np.asarray(one)[inner[1:]], ref[1][inner[1:]]) <= (1e-10 if real_t == np.float64 else 1e-5)


@pytest.mark.parametrize("real_t", [np.float64, np.float32], ids=["f64", "f32"])
def test_poisson_2d_against_oracle(cuda, real_t):
    from sopht_mpi_b200.numeric.eulerian_grid_ops import UnboundedPoissonSolverMPI2D

    gs, n = 2, (16, 32)
    rng = np.random.default_rng(2)
    shape = tuple(v + 2 * gs for v in n)
    mc, _ = _constructs(2, n, real_t)
    solver = UnboundedPoissonSolverMPI2D(*n, mpi_construct=mc, ghost_size=gs, x_range=1.0, real_t=real_t)
    oracle = UnboundedPoissonSolverOracle2D(*n, x_range=1.0, real_t=real_t)
    rhs = rng.uniform(size=shape).astype(real_t)
    ref = np.zeros_like(rhs)
    oracle.solve(ref, rhs, gs)
    out = _dev(np.zeros_like(rhs), cuda)
    solver.solve(solution_field=out, rhs_field=_dev(rhs, cuda))
    inner = (slice(gs, -gs),) * 2
    assert _rel(np.asarray(out)[inner], ref[inner]) <= (1e-10 if real_t == np.float64 else 1e-5)


IB_FILES = sorted(glob.glob(os.path.join(GOLDEN, "ib_3d_*.npz")) + glob.glob(os.path.join(GOLDEN, "ib_2d_f*_f64.npz")))


@pytest.mark.parametrize("path", [p for p in IB_FILES if "_sub" not in p],
                         ids=lambda p: os.path.basename(p))
def test_virtual_boundary_forcing_against_reference_golden(cuda, path):
    """VBF through the public class, single rank (substart 0), against vectors from the
    reference's own numba kernels."""
    from sopht_mpi_b200.numeric.immersed_boundary_ops import VirtualBoundaryForcingMPI

    gd = np.load(path)
    dim, gs = int(gd["dim"]), int(gd["gs"])
    dx = gd["dx"][()]
    real_t = type(dx)
    n_local = gd["eul_vec"].shape[-1] - 2 * gs
    mc, _ = _constructs(dim, (n_local,) * dim, real_t)
    pos = gd["pos"]
    vel = np.full_like(pos, 0.25)
    vbf = VirtualBoundaryForcingMPI(mpi_construct=mc, ghost_size=gs,
                                    virtual_boundary_stiffness_coeff=-3.0,
                                    virtual_boundary_damping_coeff=-0.5, grid_dim=dim, dx=dx,
                                    global_lag_grid_position_field=pos)
    vbf._fetch_index_and_weights = True
    vbf.local_lag_grid_position_mismatch_field[...] = 0.01
    eul_u = _dev(gd["eul_vec"], cuda)
    eul_f = _dev(np.ones_like(gd["eul_vec"]), cuda)
    vbf.compute_interaction_forcing(local_eul_grid_forcing_field=eul_f,
                                    local_eul_grid_velocity_field=eul_u,
                                    global_lag_grid_position_field=pos,
                                    global_lag_grid_velocity_field=vel)
    tol = 1e-5 if (pos.dtype == np.float32) else 1e-10
    assert np.array_equal(vbf.local_nearest_eul_grid_index_to_lag_grid, gd["nearest"])
    assert _rel(vbf.local_interp_weights, gd["w_cos"]) <= tol
    assert _rel(vbf.local_lag_grid_flow_velocity_field, gd["e2l_vec"]) <= tol
    force = -3.0 * 0.01 - 0.5 * (gd["e2l_vec"] - vel)
    assert _rel(vbf.global_lag_grid_forcing_field, force) <= 10 * tol
    # spreading of that force + ghost clearing, via the oracle spreading of the golden weights
    ref = np.zeros_like(gd["eul_vec"])
    ib_oracle.lagrangian_to_eulerian(ref, force.astype(pos.dtype), gd["w_cos"], gd["nearest"], 2)
    ib_oracle.clear_ghost_cells_nd(ref, gs, dim)
    assert _rel(eul_f, ref) <= (1e-5 if real_t == np.float32 else 1e-10)
    before = vbf.local_lag_grid_position_mismatch_field.copy()
    vbf.time_step(dt=0.1)
    assert np.allclose(vbf.local_lag_grid_position_mismatch_field,
                       before + 0.1 * vbf.local_lag_grid_velocity_mismatch_field)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ib_*_sub.npz"))),
                         ids=lambda p: os.path.basename(p))
def test_ib_kernels_with_rank_offset_against_golden(cuda, path):
    """C ABI called directly with a non-zero substart (what an inner slab sees)."""
    import ctypes

    from sopht_mpi_b200 import _lib
    from sopht_mpi_b200.utils.device import dptr

    lib = _lib.load()
    gd = np.load(path)
    dim, gs, w = int(gd["dim"]), int(gd["gs"]), int(gd["width"])
    dx, shift = gd["dx"][()], gd["shift"][()]
    real_t, pos = type(dx), gd["pos"]
    n_local = gd["eul_vec"].shape[-1] - 2 * gs
    g = _lib.make_grid(dim, real_t, gs, (n_local,) * dim, [1] * (2 * dim))
    p = _lib.IBParams()
    p.lag_dtype, p.kernel_type, p.width = _lib.dtype_code(pos.dtype), 0, w
    ss = list(gd["substart_xyz"]) + [0] * (3 - dim)
    for i in range(3):
        p.substart_xyz[i] = int(ss[i])
    p.dx, p.coord_shift = float(dx), float(shift)
    n = pos.shape[1]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)  # noqa: E731
    near = torch.zeros((dim, n), dtype=torch.int64, device=cuda)
    wts = t(np.zeros_like(gd["w_cos"]))
    u = t(np.zeros_like(pos))
    eul_d, pos_d = t(gd["eul_vec"]), t(pos)  # keep the device buffers alive across the launch
    _lib.check(lib, lib.sb200_ib_interact_lag(ctypes.byref(g), ctypes.byref(p), n, dptr(eul_d),
                                              dptr(pos_d), None, None, dptr(near), dptr(wts), dptr(u),
                                              None, None, None))
    torch.cuda.synchronize()
    tol = 1e-5 if pos.dtype == np.float32 else 1e-10
    assert np.array_equal(near.cpu().numpy(), gd["nearest"])
    assert _rel(wts.cpu().numpy(), gd["w_cos"]) <= tol
    assert _rel(u.cpu().numpy(), gd["e2l_vec"]) <= tol


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "unfused"])
@pytest.mark.parametrize("real_t,steps", [(np.float64, 5), (np.float32, 5)], ids=["f64", "f32"])
@pytest.mark.parametrize("flow_type", ["navier_stokes_with_forcing", "navier_stokes"])
def test_simulator_steps_against_oracle(cuda, real_t, steps, flow_type, fused):
    """N full time steps (same dt fed to both) on a smooth compact vorticity field +
    a smooth forcing: omega, u, psi within the north-star tolerance."""
    from sopht_mpi_b200.simulator import UnboundedFlowSimulator3D

    n = (24, 24, 32)
    kw = dict(grid_size=n, x_range=1.0, kinematic_viscosity=1e-2, flow_type=flow_type, real_t=real_t,
              with_free_stream_flow=True, filter_vorticity=True,
              filter_setting_dict={"order": 1, "type": "multiplicative"})
    sim = UnboundedFlowSimulator3D(use_fused_kernels=fused, **kw)
    ora = FlowSimulatorOracle3D(**kw)
    gs = sim.ghost_size
    x, y, z = ora.local_x, ora.local_y, ora.local_z
    zz, yy, xx = np.meshgrid(z, y, x, indexing="ij")
    r2 = (xx - 0.5) ** 2 + (yy - 0.375) ** 2 + (zz - 0.375) ** 2
    blob = np.exp(-r2 / 0.01)
    w0 = np.stack([blob * (yy - 0.375) * 40, -blob * (xx - 0.5) * 40, 0.3 * blob]).astype(real_t)
    ora.vorticity_field[...] = w0
    sim.vorticity_field[...] = w0
    u_inf = [1.0, 0.0, 0.0]
    ora.compute_flow_velocity(u_inf)
    sim.compute_flow_velocity(free_stream_velocity=u_inf)
    assert _rel(sim.velocity_field, ora.velocity_field) <= TOL[real_t]
    for step in range(steps):
        dt = ora.compute_stable_timestep(dt_prefac=0.5)
        dt_gpu = sim.compute_stable_timestep(dt_prefac=0.5)
        assert abs(dt_gpu - dt) <= 1e-5 * dt
        if flow_type == "navier_stokes_with_forcing":
            force = (0.5 * np.stack([blob, 0.5 * blob, -blob]) * np.cos(3.0 * step)).astype(real_t)
            force[:, :gs], force[:, -gs:] = 0, 0
            ora.eul_grid_forcing_field[...] = force
            sim.eul_grid_forcing_field[...] = force
        ora.time_step(dt, free_stream_velocity=u_inf)
        sim.time_step(dt=dt, free_stream_velocity=u_inf)
    tol = TOL[real_t]
    assert _rel(sim.vorticity_field, ora.vorticity_field) <= tol
    assert _rel(sim.velocity_field, ora.velocity_field) <= tol
    assert _rel(sim.stream_func_field, ora.stream_func_field) <= tol
    if flow_type == "navier_stokes_with_forcing":
        assert float(np.abs(np.asarray(sim.eul_grid_forcing_field)).max()) == 0.0
    assert abs(sim.get_max_vorticity() - ora.vorticity_field[:, gs:-gs, gs:-gs, gs:-gs].max()) <= tol * 50
    assert sim.time == pytest.approx(ora.time)


@pytest.mark.parametrize("real_t", [np.float64, np.float32], ids=["f64", "f32"])
def test_simulator_2d_cylinder_against_oracle(cuda, real_t):
    """BASELINE config 1 (2D flow past a cylinder: navier_stokes_with_forcing + free stream +
    virtual boundary forcing on 60 points), at 64 x 128, against the composed CPU oracle."""
    from oracle.simulator import FlowSimulatorOracle2D
    from sopht_mpi_b200.numeric.immersed_boundary_ops import VirtualBoundaryForcingMPI
    from sopht_mpi_b200.simulator import UnboundedFlowSimulator2D

    n = (64, 128)
    radius, u_free = 0.06, 1.0
    nu = radius * u_free / 200.0
    kw = dict(grid_size=n, x_range=1.0, kinematic_viscosity=nu, flow_type="navier_stokes_with_forcing",
              real_t=real_t, with_free_stream_flow=True)
    sim = UnboundedFlowSimulator2D(**kw)
    ora = FlowSimulatorOracle2D(**kw)
    gs = sim.ghost_size
    assert sim.position_field.shape == (2,) + tuple(v + 2 * gs for v in n)
    assert np.array_equal(sim.position_field[0][0], ora.local_x)
    theta = 2 * np.pi * np.arange(60) / 60
    pos = np.stack([2.5 * radius + radius * np.cos(theta), 0.25 + radius * np.sin(theta)])
    vel = np.zeros_like(pos)
    max_lag_dx = 2 * np.pi * radius / 60
    k, c = -5e4 * max_lag_dx, -20 * max_lag_dx
    vbf = VirtualBoundaryForcingMPI(mpi_construct=sim.mpi_construct, ghost_size=gs,
                                    virtual_boundary_stiffness_coeff=k, virtual_boundary_damping_coeff=c,
                                    grid_dim=2, dx=sim.dx, global_lag_grid_position_field=pos)
    vbf_o = ib_oracle.VirtualBoundaryForcingOracle(k, c, 2, ora.dx, real_t, np.float64, gs)
    u_inf = [u_free, 0.0]
    for c_ in range(2):
        ora.velocity_field[c_] += real_t(u_inf[c_])
    sim.velocity_field[...] = ora.velocity_field
    for _ in range(6):
        dt = ora.compute_stable_timestep()
        assert abs(sim.compute_stable_timestep() - dt) <= 1e-6 * dt
        vbf_o.compute_interaction_force_on_eul_and_lag_grid(ora.eul_grid_forcing_field, ora.velocity_field,
                                                            pos, vel)
        vbf.compute_interaction_forcing(local_eul_grid_forcing_field=sim.eul_grid_forcing_field,
                                        local_eul_grid_velocity_field=sim.velocity_field,
                                        global_lag_grid_position_field=pos,
                                        global_lag_grid_velocity_field=vel)
        vbf_o.time_step(dt)
        vbf.time_step(dt=dt)
        ora.time_step(dt, free_stream_velocity=u_inf)
        sim.time_step(dt=dt, free_stream_velocity=u_inf)
    tol = TOL[real_t]
    assert np.abs(ora.vorticity_field).max() > 1.0  # the body did shed vorticity
    assert _rel(sim.vorticity_field, ora.vorticity_field) <= tol
    assert _rel(sim.velocity_field, ora.velocity_field) <= tol
    assert _rel(sim.stream_func_field, ora.stream_func_field) <= tol
    assert _rel(vbf.global_lag_grid_forcing_field, vbf_o.forcing) <= 10 * tol
    assert abs(sim.get_max_vorticity() - ora.vorticity_field[gs:-gs, gs:-gs].max()) <= 50 * tol * np.abs(
        ora.vorticity_field).max()


@pytest.mark.parametrize("flow_type", ["passive_scalar", "navier_stokes"])
def test_simulator_2d_other_flow_types(cuda, flow_type):
    from oracle.simulator import FlowSimulatorOracle2D
    from sopht_mpi_b200.simulator import UnboundedFlowSimulator2D

    real_t, n = np.float64, (32, 64)
    kw = dict(grid_size=n, x_range=1.0, kinematic_viscosity=2e-3, flow_type=flow_type, real_t=real_t)
    sim, ora = UnboundedFlowSimulator2D(**kw), FlowSimulatorOracle2D(**kw)
    yy, xx = np.meshgrid(ora.local_y, ora.local_x, indexing="ij")
    q = np.exp(-((xx - 0.5) ** 2 + (yy - 0.25) ** 2) / 0.004).astype(real_t)
    ora.primary_scalar_field[...] = q
    sim.primary_scalar_field[...] = q
    if flow_type == "passive_scalar":
        vel = np.stack([-(yy - 0.25), xx - 0.5]).astype(real_t)
        ora.velocity_field[...] = vel
        sim.velocity_field[...] = vel
    else:
        ora.compute_velocity_from_vorticity()
        sim.compute_velocity_from_vorticity()
    for _ in range(4):
        dt = ora.compute_stable_timestep()
        assert abs(sim.compute_stable_timestep() - dt) <= 1e-12
        ora.time_step(dt)
        sim.time_step(dt)
    gs = 2
    assert _rel(np.asarray(sim.primary_scalar_field)[gs:-gs, gs:-gs],
                ora.primary_scalar_field[gs:-gs, gs:-gs]) <= 1e-10
    assert _rel(sim.velocity_field, ora.velocity_field) <= 1e-10
    with pytest.raises(ValueError):
        UnboundedFlowSimulator2D(grid_size=n, x_range=1.0, kinematic_viscosity=1e-3, flow_type="bogus")
    with pytest.raises(ValueError):
        UnboundedFlowSimulator2D(grid_size=n, x_range=1.0, kinematic_viscosity=1e-3,
                                 flow_type="passive_scalar", with_free_stream_flow=True)


@pytest.mark.parametrize("flow_type", ["passive_scalar", "passive_vector"])
def test_passive_flows_against_oracle(cuda, flow_type):
    from sopht_mpi_b200.simulator import UnboundedFlowSimulator3D

    real_t = np.float64
    n = (16, 16, 16)
    kw = dict(grid_size=n, x_range=1.0, kinematic_viscosity=5e-3, flow_type=flow_type, real_t=real_t)
    sim = UnboundedFlowSimulator3D(**kw)
    ora = FlowSimulatorOracle3D(**kw)
    rng = np.random.default_rng(4)
    vel = (rng.uniform(size=ora.velocity_field.shape) - 0.5).astype(real_t)
    ora.velocity_field[...] = vel
    sim.velocity_field[...] = vel
    name = "primary_scalar_field" if flow_type == "passive_scalar" else "primary_vector_field"
    q = rng.uniform(size=getattr(ora, name).shape).astype(real_t)
    getattr(ora, name)[...] = q
    getattr(sim, name)[...] = q
    for _ in range(3):
        dt = ora.compute_stable_timestep()
        assert abs(sim.compute_stable_timestep() - dt) <= 1e-12
        ora.time_step(dt)
        sim.time_step(dt)
    gs = 2
    inner = (Ellipsis,) + (slice(gs, -gs),) * 3
    assert _rel(np.asarray(getattr(sim, name))[inner], getattr(ora, name)[inner]) <= 1e-10


def test_poisson_stage_profiling(cuda):
    """sb200_poisson_set_profiling / _last_stage_ms: five positive per-kernel times that add up to
    about the solve, and a loud error when no profiled solve exists."""
    from sopht_mpi_b200 import _lib
    from sopht_mpi_b200.numeric.eulerian_grid_ops import UnboundedPoissonSolverMPI3D

    n, gs, real_t = (64, 64, 64), 2, np.float32
    mc, _ = _constructs(3, n, real_t)
    solver = UnboundedPoissonSolverMPI3D(*n, mpi_construct=mc, ghost_size=gs, x_range=1.0, real_t=real_t)
    assert solver.backend == "fft"
    rhs = torch.rand((3,) + tuple(v + 2 * gs for v in n), device=cuda)
    sol = torch.zeros_like(rhs)
    solver.set_profiling(True)
    with pytest.raises(_lib.SophtB200Error):
        solver.last_stage_ms()
    solver.vector_field_solve(solution_vector_field=sol, rhs_vector_field=rhs)
    ms = solver.last_stage_ms()
    assert list(ms) == list(solver.STAGE_NAMES) and all(v > 0 for v in ms.values())
    solver.set_profiling(False)
    ref = torch.zeros_like(rhs)
    solver.vector_field_solve(solution_vector_field=ref, rhs_vector_field=rhs)
    assert torch.equal(ref, sol)  # profiling does not change the result


def test_full_size_step_symmetry_and_free_stream(cuda):
    """256^3 float32 (BASELINE configs[1]): size-independent properties of the whole time step.
    A vortex ring centred in the box stays mirror-symmetric in x and y (omega_z only picks up the
    four-fold discretisation error of a 6-cell core); a uniform free stream with no vorticity is a
    fixed point of the step."""
    import bench
    from sopht_mpi_b200.simulator import UnboundedFlowSimulator3D

    n, real_t = (256, 256, 256), np.float32
    sim = UnboundedFlowSimulator3D(grid_size=n, x_range=1.0, kinematic_viscosity=1e-3, flow_type="navier_stokes",
                                   real_t=real_t)
    x = sim.local_x[None, None, :].astype(np.float64)
    y = sim.local_y[None, :, None].astype(np.float64)
    z = sim.local_z[:, None, None].astype(np.float64)
    sim.vorticity_field[...] = torch.from_numpy(bench.vortex_ring(x, y, z, real_t)).to(cuda)
    sim.compute_flow_velocity(free_stream_velocity=[0.0, 0.0, 0.0])
    for _ in range(2):
        dt = sim.compute_stable_timestep()
        assert 0 < dt < 1
        sim.time_step(dt=dt, free_stream_velocity=[0.0, 0.0, 0.0])
    w = sim.vorticity_field.tensor
    assert torch.isfinite(w).all()
    scale = w.abs().max().item()
    assert w[2].abs().max().item() <= 0.1 * scale             # axial vorticity: discretisation error only
    assert (w[1] + w[1].flip(2)).abs().max().item() <= 2e-4 * scale   # omega_y odd under x -> 1 - x
    assert (w[0] + w[0].flip(1)).abs().max().item() <= 2e-4 * scale   # omega_x odd under y -> 1 - y
    u = sim.velocity_field.tensor
    assert (u[2] - u[2].flip(2)).abs().max().item() <= 2e-4 * u.abs().max().item()  # u_z even in x

    sim2 = UnboundedFlowSimulator3D(grid_size=(64, 64, 128), x_range=1.0, kinematic_viscosity=1e-3,
                                    flow_type="navier_stokes_with_forcing", real_t=real_t,
                                    with_free_stream_flow=True)
    u_inf = [1.0, -0.5, 0.25]
    sim2.compute_flow_velocity(free_stream_velocity=u_inf)
    sim2.time_step(dt=sim2.compute_stable_timestep(), free_stream_velocity=u_inf)
    assert float(sim2.vorticity_field.tensor.abs().max()) == 0.0
    for c in range(3):
        assert torch.all(sim2.velocity_field.tensor[c] == real_t(u_inf[c]))


def test_full_size_poisson_properties(cuda):
    """256^3 float32 (BASELINE config 2 size): linearity and -lap(psi) = omega in the
    interior for a smooth compact source (size-independent checks)."""
    from sopht_mpi_b200.numeric.eulerian_grid_ops import UnboundedPoissonSolverMPI3D

    n, gs, real_t = (256, 256, 256), 2, np.float32
    mc, _ = _constructs(3, n, real_t)
    solver = UnboundedPoissonSolverMPI3D(*n, mpi_construct=mc, ghost_size=gs, x_range=1.0, real_t=real_t)
    m = n[0] + 2 * gs
    c = (torch.arange(m, device=cuda, dtype=torch.float32) - gs + 0.5) / n[0]
    zz, yy, xx = torch.meshgrid(c, c, c, indexing="ij")
    a = torch.exp(-((xx - 0.5) ** 2 + (yy - 0.5) ** 2 + (zz - 0.4) ** 2) / 0.005)
    b = torch.exp(-((xx - 0.4) ** 2 + (yy - 0.6) ** 2 + (zz - 0.5) ** 2) / 0.008)
    pa, pb, pab = torch.zeros_like(a), torch.zeros_like(a), torch.zeros_like(a)
    solver.solve(pa, a)
    solver.solve(pb, b)
    solver.solve(pab, a + 2 * b)
    scale = pab.abs().max().item()
    assert (pab - (pa + 2 * pb)).abs().max().item() / scale < 1e-5
    dx = 1.0 / n[0]
    lap = (pa[2:, 1:-1, 1:-1] + pa[:-2, 1:-1, 1:-1] + pa[1:-1, 2:, 1:-1] + pa[1:-1, :-2, 1:-1]
           + pa[1:-1, 1:-1, 2:] + pa[1:-1, 1:-1, :-2] - 6 * pa[1:-1, 1:-1, 1:-1]) / dx ** 2
    sl = slice(gs + 8, -gs - 8)
    res = (-lap[sl, sl, sl] - a[1:-1, 1:-1, 1:-1][sl, sl, sl]).abs().max().item()
    assert res < 0.05  # second-order discretisation error of the 7-point Laplacian on this blob
