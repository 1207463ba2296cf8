"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the Eulerian stencil operators.

Plain-numpy restatement of the reference's MPI stencil wrappers for ONE
(virtual) rank.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this.

Two layers, exactly like the reference:

* ``*_serial``: the serial ``sopht`` pystencils kernels.  pystencils only
  computes on ``[ks:-ks]`` of whatever (sliced) array it is handed.  Their
  arithmetic lives in the un-vendored ``sopht`` package (SophT-Simulator
  @a281b4b, pystencils 1.1), so the formulas are the published central
  difference / 7-point / ENO3 forms -> PARITY UNPINNED
  for these formulas (no golden vectors exist in the reference).
* ``*_mpi``: the wrappers that ARE in the reference tree: interior call +
  six boundary-slab calls + zeroing of the physical-boundary ring, e.g.
  ``sopht_mpi/numeric/eulerian_grid_ops/stencil_ops_3d/diffusion_flux_mpi_3d.py:49-192``.
  ``phys`` = 6 booleans (z_prev, z_next, y_prev, y_next, x_prev, x_next) saying
  whether the neighbour along that face is ``MPI.PROC_NULL``.

Arrays are C-ordered ``(z, y, x)`` / ``(3, z, y, x)`` padded with ``gs`` ghosts.
"""
import numpy as np

ALL_PHYS = (True,) * 6


# --------------------------------------------------------------------------
# serial kernels (compute on [ks:-ks] of the passed views)
# --------------------------------------------------------------------------
def _c(a):  # centre view for ks = 1
    return a[..., 1:-1, 1:-1, 1:-1]


def _sh(a, dz, dy, dx, ks=1):
    """View of ``a`` shifted by (dz, dy, dx), restricted to the [ks:-ks] cells."""
    nz, ny, nx = a.shape[-3:]
    return a[
        ...,
        ks + dz : nz - ks + dz,
        ks + dy : ny - ks + dy,
        ks + dx : nx - ks + dx,
    ]


def curl_of(field, prefactor):
    """p * curl(field) on the [1:-1] cells; field is (3, z, y, x), order x,y,z.
    SURVEY.md Appendix A (gen_curl_pyst_kernel_3d)."""
    fx, fy, fz = field[0], field[1], field[2]
    p = field.dtype.type(prefactor)
    cx = p * (_sh(fz, 0, 1, 0) - _sh(fz, 0, -1, 0) - _sh(fy, 1, 0, 0) + _sh(fy, -1, 0, 0))
    cy = p * (_sh(fx, 1, 0, 0) - _sh(fx, -1, 0, 0) - _sh(fz, 0, 0, 1) + _sh(fz, 0, 0, -1))
    cz = p * (_sh(fy, 0, 0, 1) - _sh(fy, 0, 0, -1) - _sh(fx, 0, 1, 0) + _sh(fx, 0, -1, 0))
    return cx, cy, cz


def curl_serial(curl, field, prefactor):
    if min(field.shape[1:]) < 3:
        return
    cx, cy, cz = curl_of(field, prefactor)
    _c(curl[0])[...] = cx
    _c(curl[1])[...] = cy
    _c(curl[2])[...] = cz


def update_vorticity_from_velocity_forcing_serial(
    vorticity_field, velocity_forcing_field, prefactor
):
    """omega += p * curl(F) on [1:-1] (reference docstring
    update_vorticity_from_velocity_forcing_mpi_3d.py:30-35)."""
    if min(velocity_forcing_field.shape[1:]) < 3:
        return
    cx, cy, cz = curl_of(velocity_forcing_field, prefactor)
    _c(vorticity_field[0])[...] += cx
    _c(vorticity_field[1])[...] += cy
    _c(vorticity_field[2])[...] += cz


def diffusion_flux_serial(diffusion_flux, field, prefactor):
    """flux = p * (sum of 6 neighbours - 6 centre) on [1:-1]."""
    if min(field.shape) < 3:
        return
    p = field.dtype.type(prefactor)
    _c(diffusion_flux)[...] = p * (
        _sh(field, 0, 0, 1)
        + _sh(field, 0, 0, -1)
        + _sh(field, 0, 1, 0)
        + _sh(field, 0, -1, 0)
        + _sh(field, 1, 0, 0)
        + _sh(field, -1, 0, 0)
        - field.dtype.type(6) * _sh(field, 0, 0, 0)
    )


def divergence_serial(divergence, field, inv_dx):
    """div = (0.5 * inv_dx) * central differences on [1:-1] (Appendix A)."""
    if min(field.shape[1:]) < 3:
        return
    fx, fy, fz = field[0], field[1], field[2]
    p = field.dtype.type(0.5 * inv_dx)
    _c(divergence)[...] = p * (
        _sh(fx, 0, 0, 1)
        - _sh(fx, 0, 0, -1)
        + _sh(fy, 0, 1, 0)
        - _sh(fy, 0, -1, 0)
        + _sh(fz, 1, 0, 0)
        - _sh(fz, -1, 0, 0)
    )


def laplacian_filter_axis_serial(filter_flux, field, axis):
    """flux = 0.25 * (-f(+1) - f(-1) + 2 f) along one axis, on [1:-1]^3
    (reference laplacian_filter_mpi_3d.py:62-99; pystencils uses the same ghost
    layer count in every dimension)."""
    if min(field.shape) < 3:
        return
    d = [0, 0, 0]
    d[axis] = 1
    t = field.dtype.type
    _c(filter_flux)[...] = t(0.25) * (
        -_sh(field, *d) - _sh(field, -d[0], -d[1], -d[2]) + t(2) * _sh(field, 0, 0, 0)
    )


def _eno3_face_flux(field, vel, axis, ks):
    """Upwinded ENO3 flux through the + face along ``axis`` minus the - face,
    evaluated on the [2:-2] cells."""
    t = field.dtype.type

    def s(a, k):
        d = [0, 0, 0]
        d[axis] = k
        return _sh(a, d[0], d[1], d[2], ks=ks)

    half = t(0.5)
    vp = half * (s(vel, 0) + s(vel, 1))  # + face velocity
    vm = half * (s(vel, 0) + s(vel, -1))  # - face velocity
    sixth = t(1.0 / 6.0)
    # + face: left-biased / right-biased third order reconstructions
    fl_p = sixth * (-s(field, -1) + t(5) * s(field, 0) + t(2) * s(field, 1))
    fr_p = sixth * (t(2) * s(field, 0) + t(5) * s(field, 1) - s(field, 2))
    # - face
    fl_m = sixth * (-s(field, -2) + t(5) * s(field, -1) + t(2) * s(field, 0))
    fr_m = sixth * (t(2) * s(field, -1) + t(5) * s(field, 0) - s(field, 1))
    zero = t(0)
    flux_p = np.maximum(vp, zero) * fl_p + np.minimum(vp, zero) * fr_p
    flux_m = np.maximum(vm, zero) * fl_m + np.minimum(vm, zero) * fr_m
    return flux_p - flux_m


def advection_flux_eno3_serial(advection_flux, field, velocity, inv_dx):
    """Conservative ENO3 advection flux, ks = 2 (Appendix A)."""
    ks = 2
    if min(field.shape) < 2 * ks + 1:
        return
    t = field.dtype.type
    # velocity component order x,y,z <-> array axes 2,1,0
    total = (
        _eno3_face_flux(field, velocity[0], 2, ks)
        + _eno3_face_flux(field, velocity[1], 1, ks)
        + _eno3_face_flux(field, velocity[2], 0, ks)
    )
    advection_flux[ks:-ks, ks:-ks, ks:-ks] = t(inv_dx) * total


# --------------------------------------------------------------------------
# the seven wrapper regions
# --------------------------------------------------------------------------
def seven_regions(shape, gs, ks, interior_x_full=False):
    """Slices (z, y, x) the reference wrappers hand to the serial kernel."""
    mz, my, mx = shape
    inner = slice(gs, -gs)
    full = slice(None)
    lo = slice(gs - ks, gs + 2 * ks)

    def hi(m):
        return slice(m - (gs + 2 * ks), m - (gs - ks))

    regions = [(inner, inner, full if interior_x_full else inner)]
    regions += [(lo, inner, inner), (hi(mz), inner, inner)]
    regions += [(full, lo, inner), (full, hi(my), inner)]
    regions += [(full, full, lo), (full, full, hi(mx))]
    return regions


def clear_physical_ring(field, gs, phys=ALL_PHYS, width=1):
    """Zero ``field`` on the [:gs+width] / [-gs-width:] slabs of physical faces
    (reference diffusion_flux_mpi_3d.py:161-192, curl_mpi_3d.py:163-194)."""
    w = gs + width
    if phys[4]:
        field[..., :, :, :w] = 0
    if phys[5]:
        field[..., :, :, -w:] = 0
    if phys[2]:
        field[..., :, :w, :] = 0
    if phys[3]:
        field[..., :, -w:, :] = 0
    if phys[0]:
        field[..., :w, :, :] = 0
    if phys[1]:
        field[..., -w:, :, :] = 0


def _v(region):
    return (slice(None),) + tuple(region)


def update_vorticity_from_velocity_forcing_mpi(
    vorticity_field, velocity_forcing_field, prefactor, gs
):
    """reference update_vorticity_from_velocity_forcing_mpi_3d.py:27-176."""
    for r in seven_regions(vorticity_field.shape[1:], gs, 1):
        update_vorticity_from_velocity_forcing_serial(
            vorticity_field[_v(r)], velocity_forcing_field[_v(r)], prefactor
        )


def curl_mpi(curl, field, prefactor, gs, phys=ALL_PHYS):
    """reference curl_mpi_3d.py:29-194 (interior call does not slice x, :44-48)."""
    for r in seven_regions(field.shape[1:], gs, 1, interior_x_full=True):
        curl_serial(curl[_v(r)], field[_v(r)], prefactor)
    clear_physical_ring(curl, gs, phys)


def diffusion_flux_mpi(diffusion_flux, field, prefactor, gs, phys=ALL_PHYS):
    """reference diffusion_flux_mpi_3d.py:35-192."""
    for r in seven_regions(field.shape, gs, 1):
        diffusion_flux_serial(diffusion_flux[r], field[r], prefactor)
    clear_physical_ring(diffusion_flux, gs, phys)


def diffusion_timestep_mpi(field, diffusion_flux, nu_dt_by_dx2, gs, phys=ALL_PHYS):
    """reference diffusion_timestep_mpi_3d.py:41-59 (scalar) / :62-90 (vector:
    three scalar calls sharing one flux buffer)."""
    if field.ndim == 4:
        for c in range(3):
            diffusion_timestep_mpi(field[c], diffusion_flux, nu_dt_by_dx2, gs, phys)
        return
    diffusion_flux_mpi(diffusion_flux, field, nu_dt_by_dx2, gs, phys)
    field += diffusion_flux


def divergence_mpi(divergence, field, inv_dx, gs, phys=ALL_PHYS):
    """reference divergence_mpi_3d.py:31-198."""
    for r in seven_regions(field.shape[1:], gs, 1):
        divergence_serial(divergence[r], field[_v(r)], inv_dx)
    clear_physical_ring(divergence, gs, phys)


def advection_flux_eno3_mpi(advection_flux, field, velocity, inv_dx, gs):
    """reference advection_flux_mpi_3d.py:25-193 (no ring zeroing)."""
    for r in seven_regions(field.shape, gs, 2):
        advection_flux_eno3_serial(advection_flux[r], field[r], velocity[_v(r)], inv_dx)


def advection_timestep_mpi(field, advection_flux, velocity, dt_by_dx, gs):
    """reference advection_timestep_mpi_3d.py:40-60 / :66-93."""
    if field.ndim == 4:
        for c in range(3):
            advection_timestep_mpi(field[c], advection_flux, velocity, dt_by_dx, gs)
        return
    advection_flux[...] = 0
    advection_flux_eno3_mpi(advection_flux, field, velocity, -dt_by_dx, gs)
    field += advection_flux


def elementwise_cross_product(result_field, field_1, field_2):
    """whole padded array (reference call flow_simulators_mpi_3d.py:397-401)."""
    a, b = field_1, field_2
    r0 = a[1] * b[2] - a[2] * b[1]
    r1 = a[2] * b[0] - a[0] * b[2]
    r2 = a[0] * b[1] - a[1] * b[0]
    result_field[0], result_field[1], result_field[2] = r0, r1, r2


# --------------------------------------------------------------------------
# penalise field boundary (fully in the reference tree)
# --------------------------------------------------------------------------
def penalise_field_boundary_mpi(field, width, dx, x_grid, y_grid, z_grid, gs, phys=ALL_PHYS):
    """reference penalise_field_boundary_mpi_3d.py:185-242 (scalar) and :249-267
    (vector).  ``x_grid`` etc. are the 1D padded coordinate lines
    (reference uses the 3D position field; only the line along the axis matters).
    """
    if field.ndim == 4:
        for c in range(3):
            penalise_field_boundary_mpi(field[c], width, dx, x_grid, y_grid, z_grid, gs, phys)
        return
    if width == 0:
        return
    t = field.dtype.type
    # sine_prefactor is a python/numpy scalar baked into the generated C as a
    # double literal; the expression is evaluated in real_t.
    # (numpy 1.22 scalar rules in the reference make this a float64 value)
    sine_prefactor = (np.pi / 2) / (width * float(dx))
    w = gs + width
    x0, x1 = x_grid[gs], x_grid[-(gs + 1)]
    y0, y1 = y_grid[gs], y_grid[-(gs + 1)]
    z0, z1 = z_grid[gs], z_grid[-(gs + 1)]

    def sn(arg):
        return np.sin(t(sine_prefactor) * arg).astype(field.dtype)

    if phys[4]:
        field[:, :, :w] = field[:, :, w - 1 : w]
        field[:, :, :w] *= sn(x_grid[:w] - x0)[None, None, :]
    if phys[5]:
        field[:, :, -w:] = field[:, :, -w : -w + 1]
        field[:, :, -w:] *= sn(x1 - x_grid[-w:])[None, None, :]
    if phys[2]:
        field[:, :w, :] = field[:, w - 1 : w, :]
        field[:, :w, :] *= sn(y_grid[:w] - y0)[None, :, None]
    if phys[3]:
        field[:, -w:, :] = field[:, -w : -w + 1, :]
        field[:, -w:, :] *= sn(y1 - y_grid[-w:])[None, :, None]
    if phys[0]:
        field[:w, :, :] = field[w - 1 : w, :, :]
        field[:w, :, :] *= sn(z_grid[:w] - z0)[:, None, None]
    if phys[1]:
        field[-w:, :, :] = field[-w : -w + 1, :, :]
        field[-w:, :, :] *= sn(z1 - z_grid[-w:])[:, None, None]


# --------------------------------------------------------------------------
# Laplacian filter (fully in the reference tree)
# --------------------------------------------------------------------------
def _laplacian_filter_axis_mpi(axis, field_buffer, filter_flux_buffer, gs):
    """reference laplacian_filter_mpi_3d.py:145-264."""
    for r in seven_regions(field_buffer.shape, gs, 1):
        laplacian_filter_axis_serial(filter_flux_buffer[r], field_buffer[r], axis)


def laplacian_filter_mpi(
    field, filter_flux_buffer, field_buffer, filter_order, filter_type, gs, phys=ALL_PHYS
):
    """reference laplacian_filter_mpi_3d.py:267-385 (scalar), :404-419 (vector).
    Array axis numbering: x-filter = array axis 2, y = 1, z = 0."""
    if field.ndim == 4:
        for c in range(3):
            laplacian_filter_mpi(
                field[c], filter_flux_buffer, field_buffer, filter_order, filter_type, gs, phys
            )
        return
    clear_physical_ring(filter_flux_buffer, gs, phys)
    if filter_type == "multiplicative":
        field_buffer[...] = field
        for _ in range(filter_order):
            for axis in (2, 1, 0):
                _laplacian_filter_axis_mpi(axis, field_buffer, filter_flux_buffer, gs)
                clear_physical_ring(filter_flux_buffer, gs, phys)
                field_buffer[...] = filter_flux_buffer
        field -= filter_flux_buffer
    elif filter_type == "convolution":
        for axis in (2, 1, 0):
            field_buffer[...] = field
            for _ in range(filter_order):
                _laplacian_filter_axis_mpi(axis, field_buffer, filter_flux_buffer, gs)
                clear_physical_ring(filter_flux_buffer, gs, phys)
                field_buffer[...] = filter_flux_buffer
            field -= filter_flux_buffer
    else:
        raise ValueError("Invalid filter type")
