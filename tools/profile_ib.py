"""Host-side cost breakdown of the immersed-boundary interaction (cProfile) on the sphere
workload: python tools/profile_ib.py [n_calls]"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from sopht_mpi_b200.simulator import (PrescribedForcingGrid, RigidBodyFlowInteractionMPI,  # noqa: E402
                                      UnboundedFlowSimulator3D)


def main():
    n_calls = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    grid = (256, 256, 512)
    sim = UnboundedFlowSimulator3D(grid_size=grid, x_range=1.0, kinematic_viscosity=1e-3,
                                   flow_type="navier_stokes_with_forcing", real_t=np.float32,
                                   with_free_stream_flow=True, rank_distribution=(0, 1, 1))
    dx = float(sim.dx)
    pts = bench.sphere_points((0.25, 0.5 * sim.y_range, 0.5 * sim.z_range), 0.2, dx)

    class _Body:
        pass

    inter = RigidBodyFlowInteractionMPI(
        mpi_construct=sim.mpi_construct, mpi_ghost_exchange_communicator=sim.mpi_ghost_exchange_communicator,
        rigid_body=_Body(), eul_grid_forcing_field=sim.eul_grid_forcing_field,
        eul_grid_velocity_field=sim.velocity_field, virtual_boundary_stiffness_coeff=-1.5e5,
        virtual_boundary_damping_coeff=-87.5, dx=sim.dx, grid_dim=3,
        forcing_grid_cls=lambda grid_dim, rigid_body: PrescribedForcingGrid(grid_dim, pts, max_lag_grid_dx=dx))
    for _ in range(3):
        inter()
        inter.time_step(1e-4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n_calls):
        inter()
        inter.time_step(1e-4)
    torch.cuda.synchronize()
    print(f"{pts.shape[1]} points: {(time.perf_counter() - t0) / n_calls * 1e3:.3f} ms per interaction + time_step")
    t0 = time.perf_counter()
    for _ in range(n_calls):
        inter.compute_flow_forces_and_torques()
    torch.cuda.synchronize()
    print(f"compute_flow_forces_and_torques: {(time.perf_counter() - t0) / n_calls * 1e3:.3f} ms")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n_calls):
        inter()
        inter.time_step(1e-4)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)


if __name__ == "__main__":
    main()
