"""float64 3D step timing (the tolerance-1e-10 path): python tools/bench_f64.py N"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from sopht_mpi_b200.simulator import UnboundedFlowSimulator3D  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    for real_t in (np.float64, np.float32):
        sim = UnboundedFlowSimulator3D(grid_size=(n, n, n), x_range=1.0, kinematic_viscosity=1e-3,
                                       flow_type="navier_stokes", real_t=real_t)
        x = sim.local_x[None, None, :].astype(np.float64)
        y = sim.local_y[None, :, None].astype(np.float64)
        z = sim.local_z[:, None, None].astype(np.float64)
        sim.vorticity_field[...] = torch.from_numpy(bench.vortex_ring(x, y, z, real_t)).cuda()
        sim.compute_flow_velocity(free_stream_velocity=[0.0] * 3)
        for _ in range(3):
            sim.time_step(dt=sim.compute_stable_timestep(), free_stream_velocity=[0.0] * 3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            sim.time_step(dt=sim.compute_stable_timestep(), free_stream_velocity=[0.0] * 3)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 10 * 1e3
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            sim.unbounded_poisson_solver.vector_field_solve(solution_vector_field=sim.stream_func_field,
                                                            rhs_vector_field=sim.vorticity_field)
        b.record()
        torch.cuda.synchronize()
        print(f"n={n} {np.dtype(real_t).name}: {ms:.3f} ms per step ({n ** 3 / ms / 1e3:.0f} Mcell-updates/s), "
              f"poisson {a.elapsed_time(b) / 5:.3f} ms, backend {sim.unbounded_poisson_solver.backend}", flush=True)
        del sim
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
