"""Poisson-only timing (device resident, CUDA events) + agreement of the two backends.
    python tools/bench_poisson.py 256 512 [--reps 10] [--check]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from sopht_mpi_b200.numeric.eulerian_grid_ops import UnboundedPoissonSolverMPI3D  # noqa: E402
from sopht_mpi_b200.utils import MPIConstruct3D  # noqa: E402
from sopht_mpi_b200.utils.device import DeviceField  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [256]
    reps = 10
    check = "--check" in sys.argv
    gs = 2
    for n in sizes:
        mc = MPIConstruct3D(n, n, n, real_t=np.float32, rank_distribution=(0, 1, 1))
        solver = UnboundedPoissonSolverMPI3D(n, n, n, mpi_construct=mc, ghost_size=gs, x_range=1.0,
                                             real_t=np.float32)
        g = torch.Generator(device="cuda").manual_seed(0)
        rhs = torch.rand((3, n + 2 * gs, n + 2 * gs, n + 2 * gs), device="cuda", generator=g)
        sol = torch.zeros_like(rhs)
        r, s = DeviceField(rhs), DeviceField(sol)
        for _ in range(3):
            solver.vector_field_solve(solution_vector_field=s, rhs_vector_field=r)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            solver.vector_field_solve(solution_vector_field=s, rhs_vector_field=r)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        gbs = 86 * 4 * n ** 3 / (ms * 1e-3) / 1e9
        line = f"n={n} backend={solver.backend} {ms:.3f} ms/vector-solve  {gbs:.0f} GB/s algorithmic (86 W/cell)"
        if solver.backend == "fft":
            solver.set_profiling(True)
            acc = {}
            for _ in range(reps):
                solver.vector_field_solve(solution_vector_field=s, rhs_vector_field=r)
                for k, v in solver.last_stage_ms().items():
                    acc[k] = acc.get(k, 0.0) + v / reps
            solver.set_profiling(False)
            line += "  stages[ms]: " + " ".join(f"{v:.3f}" for v in acc.values())
        if check:
            ref = UnboundedPoissonSolverMPI3D(n, n, n, mpi_construct=mc, ghost_size=gs, x_range=1.0,
                                              real_t=np.float32, backend="cufft")
            sol2 = torch.zeros_like(rhs)
            ref.vector_field_solve(solution_vector_field=DeviceField(sol2), rhs_vector_field=r)
            inner = (slice(None),) + (slice(gs, -gs),) * 3
            err = (sol[inner] - sol2[inner]).abs().max().item() / sol2[inner].abs().max().item()
            line += f"  rel-Linf vs cuFFT backend {err:.2e}"
            del ref, sol2
        print(line, flush=True)
        del solver, rhs, sol
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
