"""Time the fused vorticity kernel and the velocity kernel alone (device resident)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from sopht_mpi_b200.simulator import UnboundedFlowSimulator3D  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    sim = UnboundedFlowSimulator3D(grid_size=(n, n, n), x_range=1.0, kinematic_viscosity=1e-3,
                                   flow_type="navier_stokes", real_t=np.float32)
    sim.vorticity_field.tensor.uniform_(-1, 1)
    sim.velocity_field.tensor.uniform_(-1, 1)
    sim.stream_func_field.tensor.uniform_(-1, 1)

    def timed(label, fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        print(f"n={n} {label}: {ms * 1e3:.1f} us", flush=True)

    timed("fused vorticity update", lambda: sim._fused_vorticity_update(1e-4))
    timed("penalise", lambda: sim.penalise_field_towards_boundary(vector_field=sim.vorticity_field))
    import ctypes
    from sopht_mpi_b200.utils.device import dptr
    ctx = sim._ctx
    fs = (ctypes.c_double * 3)(0.0, 0.0, 0.0)
    timed("velocity from stream function", lambda: ctx.call(
        "sb200_velocity_from_stream_function", ctx.gref, dptr(sim.velocity_field.tensor),
        dptr(sim.stream_func_field.tensor), 0.5 * n, fs, None, dptr(sim._max_abs_vel_dev), ctx.stream()))


if __name__ == "__main__":
    main()
