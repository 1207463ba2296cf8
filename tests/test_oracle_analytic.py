"""Analytic consistency of the oracle's restatement of the serial ``sopht`` stencils.

The arithmetic of these kernels lives in the un-vendored ``sopht`` package, so no golden vector from the
reference can pin it bit for bit ("parity unpinned").  What CAN be pinned is what
the reference documents about them: the continuous operator each kernel discretises, its sign and
component-order conventions (fields are (x, y, z) components on C-ordered (z, y, x) arrays), the
prefactor each wrapper passes (``1 / (2 dx)``, ``1 / dx^2``, ``1 / dx``) and the order of accuracy
(second-order central differences, third-order upwind-biased ENO reconstruction).  A wrong sign, a
swapped axis, a wrong stencil coefficient or a wrong upwind branch fails these tests.
"""
import numpy as np
import pytest

from oracle import stencils as st


def _grid(n):
    """periodic box [0, 2 pi)^3 sampled at cell centres, padded so that every stencil sees valid data"""
    dx = 2 * np.pi / n
    pad = 3
    line = (np.arange(-pad, n + pad) + 0.5) * dx
    z, y, x = np.meshgrid(line, line, line, indexing="ij")
    return dx, pad, x, y, z


def _err(got, want, pad):
    inner = (Ellipsis,) + (slice(pad, -pad),) * 3
    return np.abs(np.asarray(got)[inner] - np.asarray(want)[inner]).max()


def _order(errors):
    return np.log2(errors[0] / errors[1])


def test_curl_is_the_right_handed_curl_in_xyz_component_order():
    errs = []
    for n in (32, 64):
        dx, pad, x, y, z = _grid(n)
        f = np.stack([np.sin(y) * np.cos(2 * z), np.sin(z) * np.cos(x), np.sin(2 * x) * np.cos(y)])  # (fx, fy, fz)
        want = np.stack([
            -np.sin(2 * x) * np.sin(y) - np.cos(z) * np.cos(x),          # d fz/dy - d fy/dz
            -2 * np.sin(y) * np.sin(2 * z) - 2 * np.cos(2 * x) * np.cos(y),  # d fx/dz - d fz/dx
            -np.sin(z) * np.sin(x) - np.cos(y) * np.cos(2 * z),          # d fy/dx - d fx/dy
        ])
        got = np.zeros_like(f)
        st.curl_serial(got, f, 1.0 / (2 * dx))  # the prefactor the simulator passes (flow_simulators_mpi_3d.py:388-393)
        errs.append(_err(got, want, pad))
    assert errs[1] < 3e-2 and 1.9 < _order(errs) < 2.1, errs


def test_divergence_and_laplacian_prefactors_and_second_order():
    e_div, e_lap = [], []
    for n in (16, 32):
        dx, pad, x, y, z = _grid(n)
        f = np.stack([np.sin(x) * np.cos(y), np.sin(2 * y) * np.cos(z), np.sin(z) * np.cos(x)])
        div = np.zeros_like(x)
        st.divergence_serial(div, f, 1.0 / dx)  # inv_dx as in get_vorticity_divergence_l2_norm
        e_div.append(_err(div, np.cos(x) * np.cos(y) + 2 * np.cos(2 * y) * np.cos(z) + np.cos(z) * np.cos(x), pad))
        g = np.sin(x) * np.sin(2 * y) * np.cos(z)
        lap = np.zeros_like(g)
        st.diffusion_flux_serial(lap, g, 1.0 / dx ** 2)  # nu dt / dx^2 with nu dt = 1
        e_lap.append(_err(lap, -6 * g, pad))
    assert 1.8 < _order(e_div) < 2.2 and 1.8 < _order(e_lap) < 2.2, (e_div, e_lap)


@pytest.mark.parametrize("sign", [1.0, -1.0], ids=["u>0", "u<0"])
def test_eno3_flux_is_the_conservative_advection_term_third_order_on_both_upwind_branches(sign):
    """advection_flux = div(u f) (what ``advection_timestep`` subtracts, times dt): third-order for smooth
    data with either sign of the velocity, exact for quadratics, and telescoping (conservative)."""
    errs = []
    for n in (16, 32):
        dx, pad, x, y, z = _grid(n)
        f = np.sin(x) * np.cos(y) + np.sin(z)
        u = sign * np.stack([1.0 + 0 * x, 0.5 + 0 * x, 2.0 + 0 * x])  # (ux, uy, uz), constant
        want = sign * (np.cos(x) * np.cos(y) - 0.5 * np.sin(x) * np.sin(y) + 2.0 * np.cos(z))
        flux = np.zeros_like(f)
        st.advection_flux_eno3_serial(flux, f, u, 1.0 / dx)
        errs.append(_err(flux, want, pad))
    assert errs[1] < 5e-3 and 2.7 < _order(errs) < 3.3, errs
    # exact for quadratic profiles (third-order reconstruction), any smooth positive / negative velocity
    dx, pad, x, y, z = _grid(8)
    q = 0.3 * x ** 2 - 0.2 * x * y + 0.1 * z ** 2 + y
    u = sign * np.stack([1.0 + 0 * x, 0.5 + 0 * x, 2.0 + 0 * x])
    flux = np.zeros_like(q)
    st.advection_flux_eno3_serial(flux, q, u, 1.0 / dx)
    want = sign * (1.0 * (0.6 * x - 0.2 * y) + 0.5 * (-0.2 * x + 1.0) + 2.0 * 0.2 * z)
    assert _err(flux, want, pad) < 1e-11
    # conservative: with a variable velocity the fluxes telescope, so the sum over a block of cells only
    # depends on the faces of the block -> shifting the interior data leaves it unchanged
    rng = np.random.default_rng(0)
    f = rng.standard_normal(x.shape)
    u = sign * (1.0 + 0.2 * rng.random((3,) + x.shape))
    a = np.zeros_like(f)
    st.advection_flux_eno3_serial(a, f, u, 1.0)
    f2 = f.copy()
    f2[6:-6, 6:-6, 6:-6] += rng.standard_normal(f2[6:-6, 6:-6, 6:-6].shape)
    b = np.zeros_like(f)
    st.advection_flux_eno3_serial(b, f2, u, 1.0)
    blk = (slice(3, -3),) * 3
    assert abs(a[blk].sum() - b[blk].sum()) < 1e-9 * np.abs(a[blk]).sum()


def test_vorticity_update_and_cross_product_conventions():
    """omega += p curl(F) (update_vorticity_from_velocity_forcing) and u x omega in (x, y, z) order"""
    dx, pad, x, y, z = _grid(16)
    a = np.stack([np.sin(y), np.cos(z), np.sin(x)])
    b = np.stack([np.cos(x), np.sin(z), np.cos(y)])
    got = np.zeros_like(a)
    st.elementwise_cross_product(got, a, b)
    assert np.allclose(got, np.cross(a, b, axis=0))
    w = np.zeros_like(a)
    st.update_vorticity_from_velocity_forcing_serial(w, a, 0.5)
    c = np.zeros_like(a)
    st.curl_serial(c, a, 0.5)
    assert np.array_equal(w, c)
