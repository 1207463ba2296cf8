"""BASELINE configs[0]: 2D flow past a cylinder, 512 x 256 float64, virtual boundary forcing on 60
points.  Prints ms per step (dt + interaction + flow step), single GPU."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from sopht_mpi_b200.numeric.immersed_boundary_ops import VirtualBoundaryForcingMPI  # noqa: E402
from sopht_mpi_b200.simulator import UnboundedFlowSimulator2D  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    real_t = np.float64
    n = (256, 512)
    radius, u_free = 0.03, 1.0
    sim = UnboundedFlowSimulator2D(grid_size=n, x_range=1.0, kinematic_viscosity=radius * u_free / 200.0,
                                   flow_type="navier_stokes_with_forcing", real_t=real_t,
                                   with_free_stream_flow=True, CFL=0.1)
    theta = 2 * np.pi * np.arange(60) / 60
    pos = np.stack([2.5 * radius + radius * np.cos(theta), 0.25 + radius * np.sin(theta)])
    vel = np.zeros_like(pos)
    max_lag_dx = 2 * np.pi * radius / 60
    vbf = VirtualBoundaryForcingMPI(mpi_construct=sim.mpi_construct, ghost_size=sim.ghost_size,
                                    virtual_boundary_stiffness_coeff=-5e4 * max_lag_dx,
                                    virtual_boundary_damping_coeff=-20 * max_lag_dx, grid_dim=2, dx=sim.dx,
                                    global_lag_grid_position_field=pos)
    sim.velocity_field[...] = np.broadcast_to(np.array([u_free, 0.0]).reshape(2, 1, 1),
                                              (2,) + tuple(v + 4 for v in n)).copy()

    def one_step():
        dt = sim.compute_stable_timestep()
        vbf.compute_interaction_forcing(local_eul_grid_forcing_field=sim.eul_grid_forcing_field,
                                        local_eul_grid_velocity_field=sim.velocity_field,
                                        global_lag_grid_position_field=pos, global_lag_grid_velocity_field=vel)
        vbf.time_step(dt=dt)
        sim.time_step(dt=dt, free_stream_velocity=[u_free, 0.0])

    for _ in range(10):
        one_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / steps * 1e3
    print(f"2D cylinder 512x256 f64, 60 forcing points: {ms:.3f} ms per step, "
          f"{n[0] * n[1] / ms / 1e3:.1f} Mcell-updates/s, max vorticity {sim.get_max_vorticity():.3f}")
    if "--profile" in sys.argv:
        import cProfile
        import pstats

        pr = cProfile.Profile()
        pr.enable()
        for _ in range(100):
            one_step()
        torch.cuda.synchronize()
        pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(22)


if __name__ == "__main__":
    main()
