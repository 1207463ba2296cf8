"""Kernel generators with the reference's names, arguments and error behaviour
(``sopht_mpi/numeric/eulerian_grid_ops/stencil_ops_{2,3}d/*.py``), backed by the
sm_100a kernels of ``libsophtb200``.

Every generator returns a closure taking the reference's keyword arguments; both
generator and closure expose ``kernel_support``.  Fields may be
:class:`~sopht_mpi_b200.utils.device.DeviceField` / torch CUDA tensors (no copies)
or numpy arrays (staged through the device, as a drop-in convenience).
"""
import ctypes

import numpy as np
import torch

from ... import _lib
from ...utils.device import Staged, current_stream_ptr, dptr
from ...utils.mpi_utils import check_valid_ghost_size_and_kernel_support


class OpContext:
    """What every operator needs: library handle, local grid descriptor, device."""

    def __init__(self, real_t, mpi_construct, ghost_exchange_communicator):
        self.lib = _lib.load()
        self.real_t = real_t
        self.mpi_construct = mpi_construct
        self.ghost_comm = ghost_exchange_communicator
        self.ghost_size = ghost_exchange_communicator.ghost_size
        self.dim = mpi_construct.grid_dim
        self.device = mpi_construct.device
        if self.device.type != "cuda":
            raise _lib.SophtB200Error("sopht_mpi_b200 operators need a CUDA device (no CPU fallback)")
        self.grid = _lib.make_grid(self.dim, real_t, self.ghost_size, mpi_construct.local_grid_size,
                                   mpi_construct.physical_faces)
        self.gref = ctypes.byref(self.grid)
        self.dtype = self.grid.dtype
        self.distributed = mpi_construct.size > 1
        self.padded_shape = tuple(int(n) + 2 * self.ghost_size for n in mpi_construct.local_grid_size)
        self.cells = int(np.prod(self.padded_shape))

    def call(self, name, *args):
        _lib.check(self.lib, getattr(self.lib, name)(*args))

    def stream(self):
        return current_stream_ptr(self.device)

    def stage(self):
        return Staged(self.device)

    def exchange_scalar(self, t):
        if self.distributed:
            self.ghost_comm.exchange_scalar_field_init(t)
            self.ghost_comm.exchange_finalise()

    def exchange_vector(self, t):
        if self.distributed:
            self.ghost_comm.exchange_vector_field_init(t)
            self.ghost_comm.exchange_finalise()


def _check_field_type(field_type):
    if field_type != "scalar" and field_type != "vector":
        raise ValueError("Invalid field type")


# ------------------------------------------------------------------ pointwise
def _gen_set_fixed_val(dim):
    def gen(real_t, field_type="scalar"):
        _check_field_type(field_type)
        lib = _lib.load()
        code = _lib.dtype_code(real_t)

        if field_type == "scalar":
            def set_fixed_val(field, fixed_val):
                st = Staged(_device_of(field))
                t = st(field, out=True)
                _lib.check(lib, lib.sb200_set_fixed_val(code, dptr(t), t.numel(), float(fixed_val),
                                                        current_stream_ptr()))
                st.finish()
            return set_fixed_val

        def vector_field_set_fixed_val(vector_field, fixed_vals):
            st = Staged(_device_of(vector_field))
            t = st(vector_field, out=True)
            for c in range(t.shape[0]):
                _lib.check(lib, lib.sb200_set_fixed_val(code, dptr(t[c]), t[c].numel(),
                                                        float(fixed_vals[c]), current_stream_ptr()))
            st.finish()
        return vector_field_set_fixed_val
    return gen


def _device_of(x):
    t = getattr(x, "tensor", x)
    if isinstance(t, torch.Tensor):
        return t.device
    return torch.device("cuda", torch.cuda.current_device())


gen_set_fixed_val_pyst_kernel_2d = _gen_set_fixed_val(2)
gen_set_fixed_val_pyst_kernel_3d = _gen_set_fixed_val(3)


def _gen_add_fixed_val(dim):
    def gen(real_t, field_type="scalar"):
        _check_field_type(field_type)
        lib = _lib.load()
        code = _lib.dtype_code(real_t)

        def add_fixed_val(sum_field, vector_field=None, fixed_vals=None, field=None, fixed_val=None):
            src = vector_field if field is None else field
            vals = fixed_vals if fixed_val is None else [fixed_val]
            st = Staged(_device_of(sum_field))
            ts, tf = st(sum_field, out=True), st(src)
            if ts.data_ptr() != tf.data_ptr():
                ts.copy_(tf)
            ncomp = 1 if field_type == "scalar" else ts.shape[0]
            arr = (ctypes.c_double * 3)(*([float(v) for v in vals] + [0.0] * (3 - len(vals))))
            _lib.check(lib, lib.sb200_add_fixed_val(code, dptr(ts), ncomp, ts.numel() // ncomp, arr,
                                                    current_stream_ptr()))
            st.finish()
        return add_fixed_val
    return gen


gen_add_fixed_val_pyst_kernel_2d = _gen_add_fixed_val(2)
gen_add_fixed_val_pyst_kernel_3d = _gen_add_fixed_val(3)


def gen_elementwise_cross_product_pyst_kernel_3d(real_t):
    lib = _lib.load()
    code = _lib.dtype_code(real_t)

    def elementwise_cross_product(result_field, field_1, field_2):
        st = Staged(_device_of(result_field))
        r, a, b = st(result_field, out=True), st(field_1), st(field_2)
        _lib.check(lib, lib.sb200_elementwise_cross_product(code, dptr(r), dptr(a), dptr(b),
                                                            r[0].numel(), current_stream_ptr()))
        st.finish()
    return elementwise_cross_product


# ------------------------------------------------------------------- stencils
def _gen_update_vorticity(dim):
    def gen(real_t, mpi_construct, ghost_exchange_communicator):
        kernel_support = 1
        gen.kernel_support = kernel_support
        check_valid_ghost_size_and_kernel_support(
            ghost_size=ghost_exchange_communicator.ghost_size, kernel_support=kernel_support)
        ctx = OpContext(real_t, mpi_construct, ghost_exchange_communicator)

        def update_vorticity_from_velocity_forcing(vorticity_field, velocity_forcing_field, prefactor):
            update_vorticity_from_velocity_forcing.kernel_support = kernel_support
            st = ctx.stage()
            w, f = st(vorticity_field, out=True), st(velocity_forcing_field, out=ctx.distributed)
            ctx.exchange_vector(f)
            ctx.call("sb200_update_vorticity_from_velocity_forcing", ctx.gref, dptr(w), dptr(f),
                     float(prefactor), ctx.stream())
            st.finish()
        update_vorticity_from_velocity_forcing.kernel_support = kernel_support
        return update_vorticity_from_velocity_forcing
    return gen


gen_update_vorticity_from_velocity_forcing_pyst_mpi_kernel_2d = _gen_update_vorticity(2)
gen_update_vorticity_from_velocity_forcing_pyst_mpi_kernel_3d = _gen_update_vorticity(3)


def gen_curl_pyst_mpi_kernel_3d(real_t, mpi_construct, ghost_exchange_communicator):
    """reference ``stencil_ops_3d/curl_mpi_3d.py:10-196``"""
    kernel_support = 1
    gen_curl_pyst_mpi_kernel_3d.kernel_support = kernel_support
    check_valid_ghost_size_and_kernel_support(
        ghost_size=ghost_exchange_communicator.ghost_size, kernel_support=kernel_support)
    ctx = OpContext(real_t, mpi_construct, ghost_exchange_communicator)

    def curl_pyst_mpi_kernel_3d(curl, field, prefactor):
        curl_pyst_mpi_kernel_3d.kernel_support = kernel_support
        st = ctx.stage()
        c, f = st(curl, out=True), st(field, out=ctx.distributed)
        ctx.exchange_vector(f)
        ctx.call("sb200_curl", ctx.gref, dptr(c), dptr(f), float(prefactor), ctx.stream())
        st.finish()
    curl_pyst_mpi_kernel_3d.kernel_support = kernel_support
    return curl_pyst_mpi_kernel_3d


def gen_outplane_field_curl_pyst_mpi_kernel_2d(real_t, mpi_construct, ghost_exchange_communicator):
    """reference ``stencil_ops_2d/outplane_field_curl_mpi_2d.py:10-141``"""
    kernel_support = 1
    gen_outplane_field_curl_pyst_mpi_kernel_2d.kernel_support = kernel_support
    check_valid_ghost_size_and_kernel_support(
        ghost_size=ghost_exchange_communicator.ghost_size, kernel_support=kernel_support)
    ctx = OpContext(real_t, mpi_construct, ghost_exchange_communicator)

    def outplane_field_curl_pyst_mpi_kernel_2d(curl, field, prefactor):
        outplane_field_curl_pyst_mpi_kernel_2d.kernel_support = kernel_support
        st = ctx.stage()
        c, f = st(curl, out=True), st(field, out=ctx.distributed)
        ctx.exchange_scalar(f)
        ctx.call("sb200_curl", ctx.gref, dptr(c), dptr(f), float(prefactor), ctx.stream())
        st.finish()
    outplane_field_curl_pyst_mpi_kernel_2d.kernel_support = kernel_support
    return outplane_field_curl_pyst_mpi_kernel_2d


def _gen_diffusion_flux(dim):
    def gen(real_t, mpi_construct, ghost_exchange_communicator, field_type="scalar"):
        _check_field_type(field_type)
        kernel_support = 1
        gen.kernel_support = kernel_support
        check_valid_ghost_size_and_kernel_support(
            ghost_size=ghost_exchange_communicator.ghost_size, kernel_support=kernel_support)
        ctx = OpContext(real_t, mpi_construct, ghost_exchange_communicator)

        def diffusion_flux(diffusion_flux, field, prefactor):
            diffusion_flux_fn.kernel_support = kernel_support
            st = ctx.stage()
            fl, f = st(diffusion_flux, out=True), st(field, out=ctx.distributed)
            ctx.exchange_scalar(f)
            ctx.call("sb200_diffusion_flux", ctx.gref, dptr(fl), dptr(f), float(prefactor), ctx.stream())
            st.finish()
        diffusion_flux_fn = diffusion_flux
        diffusion_flux_fn.kernel_support = kernel_support
        if field_type == "scalar":
            return diffusion_flux_fn

        def vector_field_diffusion_flux(vector_field_diffusion_flux, vector_field, prefactor):
            for c in range(dim):
                diffusion_flux_fn(diffusion_flux=vector_field_diffusion_flux[c], field=vector_field[c],
                                  prefactor=prefactor)
        vector_field_diffusion_flux.kernel_support = kernel_support
        return vector_field_diffusion_flux
    return gen


gen_diffusion_flux_pyst_mpi_kernel_2d = _gen_diffusion_flux(2)
gen_diffusion_flux_pyst_mpi_kernel_3d = _gen_diffusion_flux(3)


def _gen_diffusion_timestep(dim):
    def gen(real_t, mpi_construct, ghost_exchange_communicator, field_type="scalar"):
        _check_field_type(field_type)
        kernel_support = 1
        gen.kernel_support = kernel_support
        check_valid_ghost_size_and_kernel_support(
            ghost_size=ghost_exchange_communicator.ghost_size, kernel_support=kernel_support)
        ctx = OpContext(real_t, mpi_construct, ghost_exchange_communicator)

        def _run(field, diffusion_flux, nu_dt_by_dx2, ncomp):
            st = ctx.stage()
            f, fl = st(field, out=True), st(diffusion_flux, out=True)
            if ctx.distributed:
                for c in range(ncomp):
                    fc = f[c] if ncomp > 1 else f
                    ctx.exchange_scalar(fc)
                    ctx.call("sb200_diffusion_timestep", ctx.gref, dptr(fc), 1, dptr(fl),
                             float(nu_dt_by_dx2), ctx.stream())
            else:
                ctx.call("sb200_diffusion_timestep", ctx.gref, dptr(f), ncomp, dptr(fl),
                         float(nu_dt_by_dx2), ctx.stream())
            st.finish()

        if field_type == "scalar":
            def diffusion_timestep(field, diffusion_flux, nu_dt_by_dx2):
                _run(field, diffusion_flux, nu_dt_by_dx2, 1)
            diffusion_timestep.kernel_support = kernel_support
            return diffusion_timestep

        def vector_field_diffusion_timestep(vector_field, diffusion_flux, nu_dt_by_dx2):
            _run(vector_field, diffusion_flux, nu_dt_by_dx2, dim)
        vector_field_diffusion_timestep.kernel_support = kernel_support
        return vector_field_diffusion_timestep
    return gen


gen_diffusion_timestep_euler_forward_pyst_mpi_kernel_2d = _gen_diffusion_timestep(2)
gen_diffusion_timestep_euler_forward_pyst_mpi_kernel_3d = _gen_diffusion_timestep(3)


def _gen_advection_flux(dim):
    def gen(real_t, mpi_construct, ghost_exchange_communicator):
        kernel_support = 2
        gen.kernel_support = kernel_support
        check_valid_ghost_size_and_kernel_support(
            ghost_size=ghost_exchange_communicator.ghost_size, kernel_support=kernel_support)
        ctx = OpContext(real_t, mpi_construct, ghost_exchange_communicator)

        def advection_flux(advection_flux, field, velocity, inv_dx):
            st = ctx.stage()
            fl, f, v = st(advection_flux, out=True), st(field, out=ctx.distributed), st(velocity, out=ctx.distributed)
            ctx.exchange_scalar(f)
            ctx.exchange_vector(v)
            ctx.call("sb200_advection_flux_eno3", ctx.gref, dptr(fl), dptr(f), dptr(v), float(inv_dx),
                     ctx.stream())
            st.finish()
        advection_flux.kernel_support = kernel_support
        return advection_flux
    return gen


gen_advection_flux_conservative_eno3_pyst_mpi_kernel_2d = _gen_advection_flux(2)
gen_advection_flux_conservative_eno3_pyst_mpi_kernel_3d = _gen_advection_flux(3)


def _gen_advection_timestep(dim):
    def gen(real_t, mpi_construct, ghost_exchange_communicator, field_type="scalar"):
        _check_field_type(field_type)
        kernel_support = 2
        gen.kernel_support = kernel_support
        check_valid_ghost_size_and_kernel_support(
            ghost_size=ghost_exchange_communicator.ghost_size, kernel_support=kernel_support)
        ctx = OpContext(real_t, mpi_construct, ghost_exchange_communicator)

        def _run(field, advection_flux, velocity, dt_by_dx, ncomp):
            st = ctx.stage()
            f, fl, v = st(field, out=True), st(advection_flux, out=True), st(velocity, out=ctx.distributed)
            if ctx.distributed:
                for c in range(ncomp):
                    fc = f[c] if ncomp > 1 else f
                    ctx.exchange_scalar(fc)
                    ctx.exchange_vector(v)
                    ctx.call("sb200_advection_timestep_eno3", ctx.gref, dptr(fc), 1, dptr(fl), dptr(v),
                             float(dt_by_dx), ctx.stream())
            else:
                ctx.call("sb200_advection_timestep_eno3", ctx.gref, dptr(f), ncomp, dptr(fl), dptr(v),
                         float(dt_by_dx), ctx.stream())
            st.finish()

        if field_type == "scalar":
            def advection_timestep(field, advection_flux, velocity, dt_by_dx):
                _run(field, advection_flux, velocity, dt_by_dx, 1)
            advection_timestep.kernel_support = kernel_support
            return advection_timestep

        def vector_field_advection_timestep(vector_field, advection_flux, velocity, dt_by_dx):
            _run(vector_field, advection_flux, velocity, dt_by_dx, dim)
        vector_field_advection_timestep.kernel_support = kernel_support
        return vector_field_advection_timestep
    return gen


gen_advection_timestep_euler_forward_conservative_eno3_pyst_mpi_kernel_2d = _gen_advection_timestep(2)
gen_advection_timestep_euler_forward_conservative_eno3_pyst_mpi_kernel_3d = _gen_advection_timestep(3)


def gen_divergence_pyst_mpi_kernel_3d(real_t, mpi_construct, ghost_exchange_communicator):
    """reference ``stencil_ops_3d/divergence_mpi_3d.py:10-200``"""
    kernel_support = 1
    gen_divergence_pyst_mpi_kernel_3d.kernel_support = kernel_support
    check_valid_ghost_size_and_kernel_support(
        ghost_size=ghost_exchange_communicator.ghost_size, kernel_support=kernel_support)
    ctx = OpContext(real_t, mpi_construct, ghost_exchange_communicator)

    def divergence_pyst_mpi_kernel_3d(divergence, field, inv_dx):
        st = ctx.stage()
        d, f = st(divergence, out=True), st(field, out=ctx.distributed)
        ctx.exchange_vector(f)
        ctx.call("sb200_divergence", ctx.gref, dptr(d), dptr(f), float(inv_dx), ctx.stream())
        st.finish()
    divergence_pyst_mpi_kernel_3d.kernel_support = kernel_support
    return divergence_pyst_mpi_kernel_3d


def _penalise_factor_table(real_t, width, dx, gs, grid_lines):
    """Sine factors of every slab, in real_t, evaluated like the generated reference
    kernel (``penalise_field_boundary_mpi_3d.py:50-57,73-183``): front slabs use
    ``sin(pref * (x - x_start))``, back slabs ``sin(pref * (x_end - x))``."""
    t = np.dtype(real_t).type
    pref = t((np.pi / 2) / (width * float(dx)))
    w = gs + width
    rows = []
    for line in grid_lines:  # array order: (z,) y, x
        line = np.asarray(line).astype(real_t)
        start, end = line[gs], line[-(gs + 1)]
        rows.append(np.sin(pref * (line[:w] - start)).astype(real_t))
        rows.append(np.sin(pref * (end - line[-w:])).astype(real_t))
    return np.ascontiguousarray(np.stack(rows))


def _line_of(grid_field, axis_from_last):
    """1D coordinate line along one axis of a meshgrid-style position field."""
    a = getattr(grid_field, "tensor", grid_field)
    idx = [0] * a.ndim
    idx[a.ndim - 1 - axis_from_last] = slice(None)
    line = a[tuple(idx)]
    return line.cpu().numpy() if isinstance(line, torch.Tensor) else np.asarray(line)


def _gen_penalise(dim):
    def gen(width, dx, x_grid_field, y_grid_field, *args, **kwargs):
        # positional layout differs between 2D and 3D exactly as in the reference
        names = (["z_grid_field"] if dim == 3 else []) + [
            "real_t", "mpi_construct", "ghost_exchange_communicator", "field_type"]
        params = dict(zip(names, args))
        params.update(kwargs)
        field_type = params.get("field_type", "scalar")
        real_t = params["real_t"]
        mpi_construct = params["mpi_construct"]
        ghost_exchange_communicator = params["ghost_exchange_communicator"]
        if width < 0 or not isinstance(width, int):
            raise ValueError("invalid zone width")
        gen.kernel_support = 0
        ncomp = 1 if field_type == "scalar" else dim
        if width == 0:
            if field_type == "scalar":
                def penalise_field_boundary(field):
                    pass
            else:
                def penalise_field_boundary(vector_field):
                    pass
            penalise_field_boundary.kernel_support = 0
            return penalise_field_boundary
        ctx = OpContext(real_t, mpi_construct, ghost_exchange_communicator)
        lines = []
        if dim == 3:
            lines.append(_line_of(params["z_grid_field"], 2))
        lines += [_line_of(y_grid_field, 1), _line_of(x_grid_field, 0)]
        table = torch.from_numpy(
            _penalise_factor_table(real_t, width, dx, ctx.ghost_size, lines)).to(ctx.device)

        def _run(field):
            st = ctx.stage()
            f = st(field, out=True)
            ctx.call("sb200_penalise_field_boundary", ctx.gref, dptr(f), ncomp, int(width), dptr(table),
                     ctx.stream())
            st.finish()

        if field_type == "scalar":
            def penalise_field_boundary(field):
                _run(field)
        else:
            def penalise_field_boundary(vector_field):
                _run(vector_field)
        penalise_field_boundary.kernel_support = 0
        return penalise_field_boundary
    return gen


gen_penalise_field_boundary_pyst_mpi_kernel_2d = _gen_penalise(2)
gen_penalise_field_boundary_pyst_mpi_kernel_3d = _gen_penalise(3)


def gen_laplacian_filter_mpi_kernel_3d(mpi_construct, ghost_exchange_communicator, filter_order,
                                       filter_flux_buffer, field_buffer, real_t, field_type="scalar",
                                       filter_type="multiplicative",
                                       filter_flux_buffer_boundary_width=1):
    """reference ``stencil_ops_3d/laplacian_filter_mpi_3d.py:14-421``"""
    if filter_order < 0 or not isinstance(filter_order, int):
        raise ValueError("Invalid filter order")
    if filter_flux_buffer_boundary_width <= 0 or not isinstance(filter_flux_buffer_boundary_width, int):
        raise ValueError("Invalid value for filter flux buffer boundary zone")
    kernel_support = 1
    gen_laplacian_filter_mpi_kernel_3d.kernel_support = kernel_support
    check_valid_ghost_size_and_kernel_support(
        ghost_size=ghost_exchange_communicator.ghost_size, kernel_support=kernel_support)
    if filter_type not in ("multiplicative", "convolution"):
        raise ValueError("Invalid filter type")
    if field_type not in ("scalar", "vector"):
        raise ValueError("Invalid field type")
    ctx = OpContext(real_t, mpi_construct, ghost_exchange_communicator)
    type_code = 0 if filter_type == "multiplicative" else 1
    if filter_flux_buffer_boundary_width != 1:
        raise NotImplementedError("filter_flux_buffer_boundary_width != 1")

    def _scalar_distributed(f, flux, buf):
        # the chain of sb200_laplacian_filter, with a halo exchange of each stage's input
        s = ctx.stream
        ctx.call("sb200_clear_physical_ring", ctx.gref, dptr(flux), 1, 1, s())

        def chain(axes):
            src, dst = f, flux
            for k, a in enumerate(axes):
                ctx.exchange_scalar(src)
                fuse_sub = k == len(axes) - 1 and k > 0
                ctx.call("sb200_laplacian_filter_stage", ctx.gref, dptr(dst), dptr(src), a,
                         dptr(f) if fuse_sub else None, int(k == 0), s())
                src, dst = dst, (buf if dst is flux else flux)
            if len(axes) == 1:
                f.sub_(flux)
            elif src is not flux:  # the last flux belongs in filter_flux_buffer
                flux.copy_(buf)

        if filter_order == 0:
            for _ in range(1 if type_code == 0 else 3):
                f.sub_(flux)
        elif type_code == 0:
            chain([a for _ in range(filter_order) for a in (0, 1, 2)])
        else:
            for a in (0, 1, 2):
                chain([a] * filter_order)

    def _run(field, ncomp):
        st = ctx.stage()
        f = st(field, out=True)
        flux, buf = st(filter_flux_buffer, out=True), st(field_buffer, out=True)
        if ctx.distributed:
            for c in range(ncomp):
                _scalar_distributed(f[c] if ncomp > 1 else f, flux, buf)
        else:
            ctx.call("sb200_laplacian_filter", ctx.gref, dptr(f), ncomp, int(filter_order), type_code,
                     dptr(flux), dptr(buf), ctx.stream())
        st.finish()

    if field_type == "scalar":
        def scalar_field_filter(scalar_field):
            _run(scalar_field, 1)
        scalar_field_filter.kernel_support = kernel_support
        return scalar_field_filter

    def vector_field_filter(vector_field):
        _run(vector_field, 3)
    vector_field_filter.kernel_support = kernel_support
    return vector_field_filter


# ------------------------------------------------------- operators no simulator path uses
def gen_brinkmann_penalise_pyst_mpi_kernel_3d(real_t, field_type="scalar"):
    """reference ``stencil_ops_3d/brinkmann_penalise_mpi_3d.py:7-21``: pointwise (kernel support 0)
    ``penalised = (field + penalty_factor * char_func * penalty_field) / (1 + penalty_factor * char_func)``"""
    _check_field_type(field_type)
    gen_brinkmann_penalise_pyst_mpi_kernel_3d.kernel_support = 0
    lib = _lib.load()
    code = _lib.dtype_code(real_t)

    def _run(penalised, penalty_factor, char_func_field, penalty, field, ncomp):
        st = Staged(_device_of(penalised))
        out, chi, tgt, f = st(penalised, out=True), st(char_func_field), st(penalty), st(field)
        _lib.check(lib, lib.sb200_brinkmann_penalise(code, dptr(out), float(penalty_factor), dptr(chi), dptr(tgt),
                                                     dptr(f), ncomp, chi.numel(), current_stream_ptr()))
        st.finish()

    if field_type == "scalar":
        def brinkmann_penalise(penalised_field, penalty_factor, char_func_field, penalty_field, field):
            _run(penalised_field, penalty_factor, char_func_field, penalty_field, field, 1)
    else:
        def brinkmann_penalise(penalised_vector_field, penalty_factor, char_func_field, penalty_vector_field,
                               vector_field):
            _run(penalised_vector_field, penalty_factor, char_func_field, penalty_vector_field, vector_field, 3)
    brinkmann_penalise.kernel_support = 0
    return brinkmann_penalise


def gen_char_func_from_level_set_via_sine_heaviside_pyst_mpi_kernel_3d(blend_width, real_t):
    """reference ``stencil_ops_3d/char_func_from_level_set_mpi_3d.py:8-30``: 0 below ``-blend_width``, 1 above
    ``blend_width``, ``0.5 (1 + s + sin(pi s) / pi)`` with ``s = level_set / blend_width`` in between"""
    gen_char_func_from_level_set_via_sine_heaviside_pyst_mpi_kernel_3d.kernel_support = 0
    lib = _lib.load()
    code = _lib.dtype_code(real_t)

    def char_func_from_level_set_via_sine_heaviside(char_func_field, level_set_field):
        st = Staged(_device_of(char_func_field))
        chi, ls = st(char_func_field, out=True), st(level_set_field)
        _lib.check(lib, lib.sb200_char_func_from_level_set(code, dptr(chi), dptr(ls), float(blend_width),
                                                           chi.numel(), current_stream_ptr()))
        st.finish()
    char_func_from_level_set_via_sine_heaviside.kernel_support = 0
    return char_func_from_level_set_via_sine_heaviside


def gen_update_vorticity_from_penalised_velocity_pyst_mpi_kernel_3d(real_t, mpi_construct,
                                                                    ghost_exchange_communicator):
    """reference ``update_vorticity_from_velocity_forcing_mpi_3d.py:181-330``:
    ``vorticity += prefactor * curl(penalised_velocity - velocity)`` on the cells the wrapper writes"""
    kernel_support = 1
    gen_update_vorticity_from_penalised_velocity_pyst_mpi_kernel_3d.kernel_support = kernel_support
    check_valid_ghost_size_and_kernel_support(
        ghost_size=ghost_exchange_communicator.ghost_size, kernel_support=kernel_support)
    ctx = OpContext(real_t, mpi_construct, ghost_exchange_communicator)

    def update_vorticity_from_penalised_velocity(vorticity_field, penalised_velocity_field, velocity_field,
                                                 prefactor):
        st = ctx.stage()
        w = st(vorticity_field, out=True)
        up, u = st(penalised_velocity_field, out=ctx.distributed), st(velocity_field, out=ctx.distributed)
        ctx.exchange_vector(u)
        ctx.exchange_vector(up)
        ctx.call("sb200_update_vorticity_from_penalised_velocity", ctx.gref, dptr(w), dptr(up), dptr(u),
                 float(prefactor), ctx.stream())
        st.finish()
    update_vorticity_from_penalised_velocity.kernel_support = kernel_support
    return update_vorticity_from_penalised_velocity
