// Unbounded (Hockney-Eastwood) Poisson solve, generic backend: fused
// pad / Green's-multiply / crop kernels around cuFFT R2C/C2R on the doubled domain.
// Reference: numeric/eulerian_grid_ops/poisson_solver_3d/UnboundedPoissonSolverMPI3D.py:22-187,
// fft_mpi_3d.py:34-48 and the 2D twins.  Used for arbitrary grid sizes; the
// power-of-two hot path lives in poisson_fft.cu (backend 1).
#include "sb200_common.h"
#include "poisson.h"

#ifndef SB200_EMU
#include <cufft.h>

#define SB_CUFFT(call, what)                                  \
  do {                                                        \
    cufftResult r__ = (call);                                 \
    if (r__ != CUFFT_SUCCESS) {                               \
      sb_set_error("cufft %s failed: %d", what, (int)r__);    \
      return -3;                                              \
    }                                                         \
  } while (0)
#define SB_CUDA(call, what)                                                     \
  do {                                                                          \
    cudaError_t e__ = (call);                                                   \
    if (e__ != cudaSuccess) {                                                   \
      sb_set_error("%s: %s", what, cudaGetErrorString(e__));                    \
      return -2;                                                                \
    }                                                                           \
  } while (0)

struct SbCufftState {
  cufftHandle fwd = 0, inv = 0;
  void* real_buf = nullptr;   // (2nz,2ny,2nx) reals
  void* spec_buf = nullptr;   // (2nz,2ny,nx+1) complex
  void* ghat = nullptr;       // (2nz,2ny,nx+1) reals: Re(rfftn(G)) * dx^dim / prod(2n)
  size_t bytes = 0;
};

// rhs interior -> top-left corner of the (pre-zeroed) doubled buffer
template <typename T>
struct PadCopyOp {
  T* dst;
  const T* src;
  int nz, ny, nx, gs, dim;
  long long s_my, s_mx;  // padded source sizes
  SB_D void operator()(long long i) const {
    const int x = (int)(i % nx);
    const long long r = i / nx;
    const int y = (int)(r % ny), z = (int)(r / ny);
    const long long si = dim == 3 ? ((long long)(z + gs) * s_my + (y + gs)) * s_mx + (x + gs)
                                  : (long long)(y + gs) * s_mx + (x + gs);
    dst[((long long)z * 2 * ny + y) * 2 * nx + x] = src[si];
  }
};
template <typename T>
struct CropCopyOp {
  T* dst;
  const T* src;
  int nz, ny, nx, gs, dim;
  long long s_my, s_mx;
  SB_D void operator()(long long i) const {
    const int x = (int)(i % nx);
    const long long r = i / nx;
    const int y = (int)(r % ny), z = (int)(r / ny);
    const long long di = dim == 3 ? ((long long)(z + gs) * s_my + (y + gs)) * s_mx + (x + gs)
                                  : (long long)(y + gs) * s_mx + (x + gs);
    dst[di] = src[((long long)z * 2 * ny + y) * 2 * nx + x];
  }
};
template <typename T>
struct SpecMulOp {
  T* spec;  // interleaved complex
  const T* g;
  SB_D void operator()(long long i) const {
    const T s = g[i];
    spec[2 * i] *= s;
    spec[2 * i + 1] *= s;
  }
};
template <typename T>
struct GhatFromSpecOp {
  T* g;
  const T* spec;
  T scale;
  SB_D void operator()(long long i) const { g[i] = spec[2 * i] * scale; }
};

int sb_poisson_cufft_create(sb200_poisson* p, void* stream) {
  auto* st = new SbCufftState();
  p->backend_state = st;
  const long long n2z = p->dim == 3 ? 2LL * p->nz : 1, n2y = 2LL * p->ny, n2x = 2LL * p->nx;
  const long long nreal = n2z * n2y * n2x, nspec = n2z * n2y * (p->nx + 1);
  const size_t w = p->dtype == SB200_F32 ? 4 : 8;
  SB_CUDA(cudaMalloc(&st->real_buf, nreal * w), "poisson real buffer");
  SB_CUDA(cudaMalloc(&st->spec_buf, nspec * 2 * w), "poisson spectral buffer");
  SB_CUDA(cudaMalloc(&st->ghat, nspec * w), "poisson greens buffer");
  st->bytes = nreal * w + nspec * 3 * w;
  const cufftType ft = p->dtype == SB200_F32 ? CUFFT_R2C : CUFFT_D2Z;
  const cufftType it = p->dtype == SB200_F32 ? CUFFT_C2R : CUFFT_Z2D;
  if (p->dim == 3) {
    SB_CUFFT(cufftPlan3d(&st->fwd, (int)n2z, (int)n2y, (int)n2x, ft), "plan3d fwd");
    SB_CUFFT(cufftPlan3d(&st->inv, (int)n2z, (int)n2y, (int)n2x, it), "plan3d inv");
  } else {
    SB_CUFFT(cufftPlan2d(&st->fwd, (int)n2y, (int)n2x, ft), "plan2d fwd");
    SB_CUFFT(cufftPlan2d(&st->inv, (int)n2y, (int)n2x, it), "plan2d inv");
  }
  SB_CUFFT(cufftSetStream(st->fwd, (cudaStream_t)stream), "set stream");
  // Green's function on the doubled grid -> spectrum -> real scaled table
  int e = sb_poisson_fill_greens(p, st->real_buf, stream);
  if (e) return e;
  if (p->dtype == SB200_F32)
    SB_CUFFT(cufftExecR2C(st->fwd, (cufftReal*)st->real_buf, (cufftComplex*)st->spec_buf), "R2C(G)");
  else
    SB_CUFFT(cufftExecD2Z(st->fwd, (cufftDoubleReal*)st->real_buf, (cufftDoubleComplex*)st->spec_buf),
             "D2Z(G)");
  double dxp = 1.0;
  for (int d = 0; d < p->dim; ++d) dxp *= p->dx;
  const double scale = dxp / (double)nreal;
  SB_DISPATCH_DTYPE(p->dtype, e = sb_launch_flat(nspec,
                                                 GhatFromSpecOp<T>{(T*)st->ghat, (const T*)st->spec_buf,
                                                                   (T)scale},
                                                 stream, "ghat"));
  if (e) return e;
  SB_CUDA(cudaStreamSynchronize((cudaStream_t)stream), "poisson create sync");
  return 0;
}

int sb_poisson_cufft_destroy(sb200_poisson* p) {
  auto* st = (SbCufftState*)p->backend_state;
  if (!st) return 0;
  if (st->fwd) cufftDestroy(st->fwd);
  if (st->inv) cufftDestroy(st->inv);
  cudaFree(st->real_buf);
  cudaFree(st->spec_buf);
  cudaFree(st->ghat);
  delete st;
  p->backend_state = nullptr;
  return 0;
}

int64_t sb_poisson_cufft_bytes(const sb200_poisson* p) {
  auto* st = (SbCufftState*)p->backend_state;
  return st ? (int64_t)st->bytes : 0;
}

int sb_poisson_cufft_solve(sb200_poisson* p, void* solution, const void* rhs, int ncomp, void* stream) {
  auto* st = (SbCufftState*)p->backend_state;
  const long long nzz = p->dim == 3 ? p->nz : 1;
  const long long n2z = p->dim == 3 ? 2LL * p->nz : 1, n2y = 2LL * p->ny, n2x = 2LL * p->nx;
  const long long nreal = n2z * n2y * n2x, nspec = n2z * n2y * (p->nx + 1);
  const long long nint = nzz * p->ny * p->nx;
  const size_t w = p->dtype == SB200_F32 ? 4 : 8;
  const long long smy = p->ny + 2 * p->gs, smx = p->nx + 2 * p->gs;
  const long long svol = (p->dim == 3 ? (p->nz + 2LL * p->gs) : 1) * smy * smx;
  SB_CUFFT(cufftSetStream(st->fwd, (cudaStream_t)stream), "set stream");
  SB_CUFFT(cufftSetStream(st->inv, (cudaStream_t)stream), "set stream");
  for (int c = 0; c < ncomp; ++c) {
    const char* rc = (const char*)rhs + (size_t)c * svol * w;
    char* sc = (char*)solution + (size_t)c * svol * w;
    int e = sb_memset_async(st->real_buf, 0, nreal * w, stream);
    SB_REQUIRE(e == 0, "poisson memset failed");
    SB_DISPATCH_DTYPE(p->dtype, e = sb_launch_flat(nint,
                                                   PadCopyOp<T>{(T*)st->real_buf, (const T*)rc, p->nz,
                                                                p->ny, p->nx, p->gs, p->dim, smy, smx},
                                                   stream, "pad_copy"));
    if (e) return e;
    if (p->dtype == SB200_F32)
      SB_CUFFT(cufftExecR2C(st->fwd, (cufftReal*)st->real_buf, (cufftComplex*)st->spec_buf), "R2C");
    else
      SB_CUFFT(cufftExecD2Z(st->fwd, (cufftDoubleReal*)st->real_buf, (cufftDoubleComplex*)st->spec_buf),
               "D2Z");
    SB_DISPATCH_DTYPE(p->dtype,
                      e = sb_launch_flat(nspec, SpecMulOp<T>{(T*)st->spec_buf, (const T*)st->ghat}, stream,
                                         "spec_mul"));
    if (e) return e;
    if (p->dtype == SB200_F32)
      SB_CUFFT(cufftExecC2R(st->inv, (cufftComplex*)st->spec_buf, (cufftReal*)st->real_buf), "C2R");
    else
      SB_CUFFT(cufftExecZ2D(st->inv, (cufftDoubleComplex*)st->spec_buf, (cufftDoubleReal*)st->real_buf),
               "Z2D");
    SB_DISPATCH_DTYPE(p->dtype, e = sb_launch_flat(nint,
                                                   CropCopyOp<T>{(T*)sc, (const T*)st->real_buf, p->nz,
                                                                 p->ny, p->nx, p->gs, p->dim, smy, smx},
                                                   stream, "crop_copy"));
    if (e) return e;
  }
  return 0;
}
#else
int sb_poisson_cufft_create(sb200_poisson*, void*) { sb_set_error("cuFFT backend unavailable in emulation"); return -1; }
int sb_poisson_cufft_destroy(sb200_poisson*) { return 0; }
int64_t sb_poisson_cufft_bytes(const sb200_poisson*) { return 0; }
int sb_poisson_cufft_solve(sb200_poisson*, void*, const void*, int, void*) { return -1; }
#endif
