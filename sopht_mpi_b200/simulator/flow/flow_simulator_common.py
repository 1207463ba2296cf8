"""What the 2D and 3D flow simulators share: the local (ghosted) coordinate lines, time accounting
and the device-side reductions.  The public attributes are the reference's
(``flow_simulators_mpi_{2,3}d.py``): ``dx``, ``x_range`` / ``y_range`` / ``z_range``, ``time``,
``local_grid_size_with_ghost``; the coordinate lines ``local_x`` / ``local_y`` / ``local_z`` are this
package's (the reference only keeps the meshgrid)."""
import numpy as np

from ...utils import logger
from ...utils.device import dptr

_AXIS_NAMES = ("x", "y", "z")


class FlowSimulatorCommon:
    grid_dim: int

    def _init_local_coordinates(self):
        """Cell-centre coordinates of this rank's padded block, one line per axis.

        Every axis has the spacing ``dx = real_t(x_range / grid_size_x)``; the first interior cell of
        the global grid sits at ``dx / 2``.  A rank that starts at global index ``s`` and owns ``n``
        cells gets ``n + 2 ghost_size`` points from ``(s - ghost_size + 1/2) dx`` to
        ``(s + n + ghost_size - 1/2) dx``, evaluated like the reference does (linspace in the
        precision of ``dx``, then a cast to ``real_t``) so that penalisation factors and Lagrangian
        support indices see identical coordinates."""
        dim = self.grid_dim
        size = tuple(int(v) for v in self.grid_size)  # array order: (z,) y, x
        self.dx = self.real_t(self.x_range / size[-1])
        for k in range(1, dim):  # y_range, z_range follow from the aspect ratio
            setattr(self, f"{_AXIS_NAMES[k]}_range", self.x_range * size[dim - 1 - k] / size[-1])
        local = self.mpi_construct.local_grid_size
        start = self.mpi_construct.grid.coords * local
        gs = self.ghost_size
        half, pad = self.dx / 2.0, gs * self.dx
        for axis in range(dim):  # array axis -> coordinate name (last array axis is x)
            lo, hi = start[axis] * self.dx, (start[axis] + local[axis]) * self.dx
            line = np.linspace(half + lo - pad, hi - half + pad, int(local[axis]) + 2 * gs).astype(self.real_t)
            setattr(self, f"local_{_AXIS_NAMES[dim - 1 - axis]}", line)
        self.local_grid_size_with_ghost = local + 2 * gs
        extent = "".join(f"\n{_AXIS_NAMES[k].upper()} axis from 0.0 to {getattr(self, _AXIS_NAMES[k] + '_range')}"
                         for k in range(dim))
        logger.info(f"{dim}D flow domain ready:{extent}\nbodies must be placed inside these bounds")

    # ------------------------------------------------------------------ time
    def update_simulator_time(self, dt):
        self.time += dt

    def time_step(self, dt, **kwargs):
        """One step of the configured flow type, then the clock."""
        self.flow_time_step(dt=dt, **kwargs)
        self.update_simulator_time(dt=dt)

    # ------------------------------------------------------------------ reductions
    def _reduce(self, name, field, ncomp):
        """Interior reduction ``name`` (``sb200_max_abs_sum`` / ``sb200_max`` / ``sb200_sum_squares``)
        of a device field; returns the host value."""
        ctx = self._ctx
        ctx.call(name, ctx.gref, dptr(field.tensor), ncomp, dptr(self._reduce_dev), ctx.stream())
        return float(self._reduce_dev.item())
