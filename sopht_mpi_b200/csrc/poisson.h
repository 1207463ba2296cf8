// Internal definition of the opaque Poisson handle.
#pragma once
#include "sb200_common.h"

struct sb200_poisson {
  int dim, dtype;
  int nz, ny, nx;  // global interior grid (2D: nz = 1)
  int gs;
  int rank, nranks;
  int backend;     // 0 = cuFFT generic, 1 = pruned in-kernel FFT pipeline (power-of-two grids)
  double x_range, y_range, z_range;
  double dx;       // value of real_t(x_range / nx)
  void* backend_state;
};

// Fill `dst` (device, (2nz,2ny,2nx) reals, contiguous) with the free-space Green's
// function on the doubled grid, evaluated in real_t exactly like the reference
// (UnboundedPoissonSolverMPI3D.py:82-114 / UnboundedPoissonSolverMPI2D.py:73-101).
int sb_poisson_fill_greens(const sb200_poisson* p, void* dst, void* stream);

int sb_poisson_cufft_create(sb200_poisson* p, void* stream);
int sb_poisson_cufft_destroy(sb200_poisson* p);
int sb_poisson_cufft_solve(sb200_poisson* p, void* solution, const void* rhs, int ncomp, void* stream);
int64_t sb_poisson_cufft_bytes(const sb200_poisson* p);

int sb_poisson_fft_create(sb200_poisson* p, void* stream);
int sb_poisson_fft_destroy(sb200_poisson* p);
int sb_poisson_fft_solve(sb200_poisson* p, void* solution, const void* rhs, int ncomp, void* stream);
int64_t sb_poisson_fft_bytes(const sb200_poisson* p);
