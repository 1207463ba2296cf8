"""The oracle against the golden vectors generated from the reference's own numba
kernels (tests/golden/make_golden.py) and against closed-form properties."""
import glob
import os

import numpy as np
import pytest

from oracle import ib, stencils as st
from oracle.poisson import UnboundedPoissonSolverOracle2D, UnboundedPoissonSolverOracle3D

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
IB_FILES = sorted(glob.glob(os.path.join(GOLDEN, "ib_*.npz")))


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("path", IB_FILES, ids=[os.path.basename(p) for p in IB_FILES])
def test_ib_oracle_matches_reference_numba_kernels(path):
    g = np.load(path)
    dim, gs, w = int(g["dim"]), int(g["gs"]), int(g["width"])
    dx, shift = g["dx"][()], g["shift"][()]
    real_t = type(dx)
    tol = 5e-7 if g["pos"].dtype == np.float32 else 1e-14
    near, sup = ib.support_and_nearest_index(g["pos"], dx, shift, w, g["substart_xyz"], gs)
    assert np.array_equal(near, g["nearest"])  # integer work: bit exact
    assert np.array_equal(sup, g["support"])
    assert _rel(ib.cosine_weights(sup.copy(), dx, real_t), g["w_cos"]) < tol
    assert _rel(ib.peskin_weights(sup.copy(), dx, real_t), g["w_pes"]) < tol
    lag = np.zeros_like(g["e2l_vec"])
    ib.eulerian_to_lagrangian(lag, g["eul_vec"], g["w_cos"], near, dx, w)
    assert _rel(lag, g["e2l_vec"]) < tol
    lag_s = np.zeros_like(g["e2l_scalar"])
    ib.eulerian_to_lagrangian(lag_s, g["eul_vec"][0], g["w_cos"], near, dx, w)
    assert _rel(lag_s, g["e2l_scalar"]) < tol
    eul = np.zeros_like(g["l2e_vec"])
    ib.lagrangian_to_eulerian(eul, g["lag_vec"], g["w_cos"], near, w)
    assert _rel(eul, g["l2e_vec"]) < (5e-7 if eul.dtype == np.float32 else 1e-14)


def test_rank_ownership_matches_reference():
    o = np.load(os.path.join(GOLDEN, "ownership.npz"))
    for t in ("f64", "f32"):
        a3 = ib.lag_nodes_rank_address(o["pos3_" + t], o["dx"][()], o["shift"][()], o["local3"], o["topo3"])
        a2 = ib.lag_nodes_rank_address(o["pos2_" + t], o["dx"][()], o["shift"][()], o["local2"], o["topo2"])
        assert np.array_equal(a3, o["addr3_" + t])
        assert np.array_equal(a2, o["addr2_" + t])


def test_cosine_weights_partition_of_unity():
    rng = np.random.default_rng(3)
    dx = np.float64(1 / 16)
    pos = rng.uniform(0.3, 0.6, size=(3, 5))
    near, sup = ib.support_and_nearest_index(pos, dx, dx / 2, 2, (0, 0, 0), 2)
    w = ib.cosine_weights(sup, dx, np.float64)
    assert np.allclose(w.sum(axis=(0, 1, 2)) * dx ** 3, 1.0, atol=1e-12)


@pytest.mark.parametrize("real_t,tol", [(np.float64, 3e-2), (np.float32, 3e-2)])
def test_poisson_oracle_solves_minus_laplacian(real_t, tol):
    n, gs = 32, 2
    s = UnboundedPoissonSolverOracle3D(n, n, n, x_range=1.0, real_t=real_t)
    dx = float(s.dx)
    c = (np.arange(n) + 0.5) * dx
    z, y, x = np.meshgrid(c, c, c, indexing="ij")
    sigma = 0.08
    r2 = (x - 0.5) ** 2 + (y - 0.5) ** 2 + (z - 0.5) ** 2
    rhs = np.zeros((n + 2 * gs,) * 3, dtype=real_t)
    rhs[gs:-gs, gs:-gs, gs:-gs] = np.exp(-r2 / (2 * sigma ** 2))
    psi = np.zeros_like(rhs)
    s.solve(psi, rhs, gs)
    # free-space solution of -lap psi = gaussian: sigma^2 * sqrt(pi/2)*sigma/r * erf(r/(sqrt2 sigma))
    from scipy.special import erf
    r = np.sqrt(r2)
    exact = sigma ** 3 * np.sqrt(np.pi / 2) * erf(r / (np.sqrt(2) * sigma)) / r
    got = psi[gs:-gs, gs:-gs, gs:-gs]
    assert np.abs(got - exact).max() / exact.max() < tol
    # ghosts untouched
    assert psi[0].max() == 0 and psi[:, :, -1].max() == 0


def test_poisson_oracle_2d_runs_and_is_linear():
    n, gs = 16, 2
    s = UnboundedPoissonSolverOracle2D(n, 2 * n, x_range=1.0, real_t=np.float64)
    rng = np.random.default_rng(0)
    a = rng.uniform(size=(n + 2 * gs, 2 * n + 2 * gs))
    b = rng.uniform(size=a.shape)
    pa, pb, pab = np.zeros_like(a), np.zeros_like(a), np.zeros_like(a)
    s.solve(pa, a, gs)
    s.solve(pb, b, gs)
    s.solve(pab, a + 2 * b, gs)
    assert np.allclose(pab, pa + 2 * pb, atol=1e-12)


def test_seven_regions_are_disjoint_and_cover_interior():
    gs, ks = 2, 1
    shape = (9, 10, 11)
    count = np.zeros(shape, dtype=int)
    for r in st.seven_regions(shape, gs, ks):
        v = count[r]
        v[ks:-ks, ks:-ks, ks:-ks] += 1
    assert count.max() == 1
    assert count[gs:-gs, gs:-gs, gs:-gs].min() == 1


def test_penalise_oracle_edge_order():
    rng = np.random.default_rng(1)
    n, gs, width = 8, 2, 2
    dx = np.float64(1 / n)
    line = np.linspace(dx / 2 - gs * dx, 1 - dx / 2 + gs * dx, n + 2 * gs)
    f = rng.uniform(size=(n + 2 * gs,) * 3)
    g = f.copy()
    st.penalise_field_boundary_mpi(g, width, dx, line, line, line, gs)
    # the first interior plane is multiplied by sin(0) = 0 ; deep interior untouched
    assert np.all(g[:, :, gs] == 0) and np.all(g[gs] == 0)
    inner = (slice(gs + width, -(gs + width)),) * 3
    assert np.array_equal(g[inner], f[inner])
    # corner value = source * sx * sy * sz
    w = gs + width
    s = np.sin((np.pi / 2) / (width * dx) * (line[:w] - line[gs]))
    assert np.isclose(g[0, 1, 3], f[w - 1, w - 1, w - 1] * s[3] * s[1] * s[0])
