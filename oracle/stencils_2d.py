"""CPU oracle (TEST INFRASTRUCTURE ONLY): 2D twins of ``oracle.stencils``.

Plain-numpy restatement of the reference's 2D MPI stencil wrappers for ONE (virtual)
rank: interior call + four boundary-slab calls (+ zeroing of the physical ring), e.g.
``sopht_mpi/numeric/eulerian_grid_ops/stencil_ops_2d/diffusion_flux_mpi_2d.py:44-140``.
The serial ``sopht`` formulas (5-point Laplacian, out-of-plane curl, ENO3) are the
published forms -> PARITY UNPINNED for the formulas, exactly as
in 3D; the wrapper regions, ring zeroing and penalisation are restated from in-tree code.

Arrays are C-ordered ``(y, x)`` / ``(2, y, x)`` padded with ``gs`` ghosts; vector
components are ordered x, y.  ``phys`` = (y_prev, y_next, x_prev, x_next).
"""
import numpy as np

ALL_PHYS = (True,) * 4


def _sh(a, dy, dx, ks=1):
    ny, nx = a.shape[-2:]
    return a[..., ks + dy : ny - ks + dy, ks + dx : nx - ks + dx]


def _c(a, ks=1):
    return a[..., ks:-ks, ks:-ks]


# ---------------------------------------------------------------- serial kernels
def diffusion_flux_serial(diffusion_flux, field, prefactor):
    if min(field.shape) < 3:
        return
    t = field.dtype.type
    _c(diffusion_flux)[...] = t(prefactor) * (
        _sh(field, 0, 1) + _sh(field, 0, -1) + _sh(field, 1, 0) + _sh(field, -1, 0)
        - t(4) * _sh(field, 0, 0)
    )


def outplane_field_curl_serial(curl, field, prefactor):
    """u_x = p (psi[y+1] - psi[y-1]); u_y = -p (psi[x+1] - psi[x-1]) on [1:-1]."""
    if min(field.shape) < 3:
        return
    p = field.dtype.type(prefactor)
    _c(curl[0])[...] = p * (_sh(field, 1, 0) - _sh(field, -1, 0))
    _c(curl[1])[...] = -p * (_sh(field, 0, 1) - _sh(field, 0, -1))


def update_vorticity_from_velocity_forcing_serial(vorticity_field, velocity_forcing_field, prefactor):
    """omega += p (dFy/dx - dFx/dy) on [1:-1]
    (reference docstring update_vorticity_from_velocity_forcing_mpi_2d.py:28-33)."""
    if min(vorticity_field.shape) < 3:
        return
    fx, fy = velocity_forcing_field[0], velocity_forcing_field[1]
    p = vorticity_field.dtype.type(prefactor)
    _c(vorticity_field)[...] += p * (_sh(fy, 0, 1) - _sh(fy, 0, -1) - _sh(fx, 1, 0) + _sh(fx, -1, 0))


def _eno3_face_flux(field, vel, axis, ks):
    t = field.dtype.type

    def s(a, k):
        d = [0, 0]
        d[axis] = k
        return _sh(a, d[0], d[1], ks=ks)

    half, sixth, zero = t(0.5), t(1.0 / 6.0), t(0)
    vp = half * (s(vel, 0) + s(vel, 1))
    vm = half * (s(vel, 0) + s(vel, -1))
    fl_p = sixth * (-s(field, -1) + t(5) * s(field, 0) + t(2) * s(field, 1))
    fr_p = sixth * (t(2) * s(field, 0) + t(5) * s(field, 1) - s(field, 2))
    fl_m = sixth * (-s(field, -2) + t(5) * s(field, -1) + t(2) * s(field, 0))
    fr_m = sixth * (t(2) * s(field, -1) + t(5) * s(field, 0) - s(field, 1))
    flux_p = np.maximum(vp, zero) * fl_p + np.minimum(vp, zero) * fr_p
    flux_m = np.maximum(vm, zero) * fl_m + np.minimum(vm, zero) * fr_m
    return flux_p - flux_m


def advection_flux_eno3_serial(advection_flux, field, velocity, inv_dx):
    ks = 2
    if min(field.shape) < 2 * ks + 1:
        return
    total = _eno3_face_flux(field, velocity[0], 1, ks) + _eno3_face_flux(field, velocity[1], 0, ks)
    advection_flux[ks:-ks, ks:-ks] = field.dtype.type(inv_dx) * total


# ---------------------------------------------------------------- wrapper regions
def five_regions(shape, gs, ks):
    """Slices (y, x) the 2D wrappers hand to the serial kernel
    (reference diffusion_flux_mpi_2d.py:49-113)."""
    my, mx = shape
    inner, full = slice(gs, -gs), slice(None)
    lo = slice(gs - ks, gs + 2 * ks)

    def hi(m):
        return slice(m - (gs + 2 * ks), m - (gs - ks))

    return [(inner, inner), (lo, inner), (hi(my), inner), (full, lo), (full, hi(mx))]


def clear_physical_ring(field, gs, phys=ALL_PHYS, width=1):
    """reference diffusion_flux_mpi_2d.py:116-140"""
    w = gs + width
    if phys[2]:
        field[..., :, :w] = 0
    if phys[3]:
        field[..., :, -w:] = 0
    if phys[0]:
        field[..., :w, :] = 0
    if phys[1]:
        field[..., -w:, :] = 0


def _v(r):
    return (slice(None),) + tuple(r)


def update_vorticity_from_velocity_forcing_mpi(vorticity_field, velocity_forcing_field, prefactor, gs):
    """reference update_vorticity_from_velocity_forcing_mpi_2d.py:25-121"""
    for r in five_regions(vorticity_field.shape, gs, 1):
        update_vorticity_from_velocity_forcing_serial(vorticity_field[r], velocity_forcing_field[_v(r)],
                                                      prefactor)


def outplane_field_curl_mpi(curl, field, prefactor, gs, phys=ALL_PHYS):
    """reference outplane_field_curl_mpi_2d.py:33-141"""
    for r in five_regions(field.shape, gs, 1):
        outplane_field_curl_serial(curl[_v(r)], field[r], prefactor)
    clear_physical_ring(curl, gs, phys)


def diffusion_flux_mpi(diffusion_flux, field, prefactor, gs, phys=ALL_PHYS):
    """reference diffusion_flux_mpi_2d.py:33-140"""
    for r in five_regions(field.shape, gs, 1):
        diffusion_flux_serial(diffusion_flux[r], field[r], prefactor)
    clear_physical_ring(diffusion_flux, gs, phys)


def diffusion_timestep_mpi(field, diffusion_flux, nu_dt_by_dx2, gs, phys=ALL_PHYS):
    """reference diffusion_timestep_mpi_2d.py:35-56"""
    diffusion_flux_mpi(diffusion_flux, field, nu_dt_by_dx2, gs, phys)
    field += diffusion_flux


def advection_timestep_mpi(field, advection_flux, velocity, dt_by_dx, gs):
    """reference advection_timestep_mpi_2d.py:36-58 + advection_flux_mpi_2d.py:24-131"""
    advection_flux[...] = 0
    for r in five_regions(field.shape, gs, 2):
        advection_flux_eno3_serial(advection_flux[r], field[r], velocity[_v(r)], -dt_by_dx)
    field += advection_flux


def penalise_field_boundary_mpi(field, width, dx, x_grid, y_grid, gs, phys=ALL_PHYS):
    """reference penalise_field_boundary_mpi_2d.py:36-170 (x faces first, then y faces)."""
    if width == 0:
        return
    t = field.dtype.type
    sine_prefactor = (np.pi / 2) / (width * float(dx))
    w = gs + width
    x0, x1 = x_grid[gs], x_grid[-(gs + 1)]
    y0, y1 = y_grid[gs], y_grid[-(gs + 1)]

    def sn(arg):
        return np.sin(t(sine_prefactor) * arg).astype(field.dtype)

    if phys[2]:
        field[:, :w] = field[:, w - 1 : w]
        field[:, :w] *= sn(x_grid[:w] - x0)[None, :]
    if phys[3]:
        field[:, -w:] = field[:, -w : -w + 1]
        field[:, -w:] *= sn(x1 - x_grid[-w:])[None, :]
    if phys[0]:
        field[:w, :] = field[w - 1 : w, :]
        field[:w, :] *= sn(y_grid[:w] - y0)[:, None]
    if phys[1]:
        field[-w:, :] = field[-w : -w + 1, :]
        field[-w:, :] *= sn(y1 - y_grid[-w:])[:, None]
