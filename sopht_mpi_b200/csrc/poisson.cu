// Poisson handle management, backend dispatch and the Green's function kernel.
#include "poisson.h"

#include <vector>

#ifndef SB200_EMU
#define SB_RN_MUL_F(a, b) __fmul_rn(a, b)
#define SB_RN_ADD_F(a, b) __fadd_rn(a, b)
#define SB_RN_MUL_D(a, b) __dmul_rn(a, b)
#define SB_RN_ADD_D(a, b) __dadd_rn(a, b)
#else
#define SB_RN_MUL_F(a, b) ((a) * (b))
#define SB_RN_ADD_F(a, b) ((a) + (b))
#define SB_RN_MUL_D(a, b) ((a) * (b))
#define SB_RN_ADD_D(a, b) ((a) + (b))
#endif
// individually rounded ops (no FMA contraction) so G matches numpy bit for bit
SB_D float rn_mul(float a, float b) { return SB_RN_MUL_F(a, b); }
SB_D float rn_add(float a, float b) { return SB_RN_ADD_F(a, b); }
SB_D double rn_mul(double a, double b) { return SB_RN_MUL_D(a, b); }
SB_D double rn_add(double a, double b) { return SB_RN_ADD_D(a, b); }

template <typename T>
struct GreensOp {
  T* dst;
  const T* xl;
  const T* yl;
  const T* zl;
  int dim;
  long long n2y, n2x;
  T two_xr, two_yr, two_zr;
  T four_pi, two_pi;
  T g0;
  SB_D void operator()(long long i) const {
    const long long x = i % n2x;
    const long long r = i / n2x;
    const long long y = r % n2y, z = r / n2y;
    const T xv = xl[x], yv = yl[y];
    const T ex = fmin(xv, two_xr - xv), ey = fmin(yv, two_yr - yv);
    T r2 = rn_add(rn_mul(ex, ex), rn_mul(ey, ey));
    T g;
    if (dim == 3) {
      const T zv = zl[z];
      const T ez = fmin(zv, two_zr - zv);
      r2 = rn_add(r2, rn_mul(ez, ez));
      g = (T(1) / sqrt(r2)) / four_pi;
    } else {
      g = -log(sqrt(r2)) / two_pi;
    }
    dst[i] = i == 0 ? g0 : g;
  }
};

// numpy.linspace(0, (n2-1)*dx, n2).astype(real_t) with dx a real_t scalar
static void sb_linspace_line(std::vector<double>& out, int n2, double dx) {
  out.resize(n2);
  const double stop = (double)(n2 - 1) * dx;
  const double step = stop / (double)(n2 - 1);
  for (int i = 0; i < n2; ++i) out[i] = (double)i * step;
  out[n2 - 1] = stop;
}

template <typename T>
static int fill_greens_t(const sb200_poisson* p, void* dst, void* stream) {
  const int n2z = p->dim == 3 ? 2 * p->nz : 1, n2y = 2 * p->ny, n2x = 2 * p->nx;
  std::vector<double> lx, ly, lz;
  sb_linspace_line(lx, n2x, p->dx);
  sb_linspace_line(ly, n2y, p->dx);
  if (p->dim == 3) sb_linspace_line(lz, n2z, p->dx); else lz.assign(1, 0.0);
  std::vector<T> h(lx.size() + ly.size() + lz.size());
  size_t k = 0;
  for (double v : lx) h[k++] = (T)v;
  for (double v : ly) h[k++] = (T)v;
  for (double v : lz) h[k++] = (T)v;
  T* d = nullptr;
#ifndef SB200_EMU
  if (cudaMalloc(&d, h.size() * sizeof(T)) != cudaSuccess) { sb_set_error("greens: cudaMalloc"); return -2; }
  cudaMemcpyAsync(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, (cudaStream_t)stream);
#else
  d = h.data();
#endif
  GreensOp<T> op;
  op.dst = (T*)dst;
  op.xl = d;
  op.yl = d + lx.size();
  op.zl = d + lx.size() + ly.size();
  op.dim = p->dim;
  op.n2y = n2y;
  op.n2x = n2x;
  op.two_xr = (T)(2 * p->x_range);
  op.two_yr = (T)(2 * p->y_range);
  op.two_zr = (T)(2 * p->z_range);
  op.four_pi = (T)(4 * M_PI);
  op.two_pi = (T)(2 * M_PI);
  if (p->dim == 3)
    op.g0 = (T)(1.0 / (4 * M_PI * p->dx));
  else
    op.g0 = (T)(-(2 * std::log(p->dx / std::sqrt(M_PI)) - 1) / (4 * M_PI));
  int e = sb_launch_flat((long long)n2z * n2y * n2x, op, stream, "greens");
#ifndef SB200_EMU
  cudaStreamSynchronize((cudaStream_t)stream);
  cudaFree(d);
#endif
  return e;
}

int sb_poisson_fill_greens(const sb200_poisson* p, void* dst, void* stream) {
  SB_DISPATCH_DTYPE(p->dtype, return fill_greens_t<T>(p, dst, stream));
}

extern "C" int sb200_poisson_create(sb200_poisson_t** out, int dim, int dtype, int nz, int ny, int nx,
                                    int gs, double x_range, int rank, int nranks, int backend,
                                    void* stream) {
  SB_REQUIRE(out, "poisson_create: null out");
  SB_REQUIRE(dim == 2 || dim == 3, "poisson_create: dim must be 2 or 3");
  SB_REQUIRE(dtype == SB200_F32 || dtype == SB200_F64, "poisson_create: bad dtype");
  SB_REQUIRE(ny > 0 && nx > 0 && (dim == 2 || nz > 0), "poisson_create: bad grid");
  auto* p = new sb200_poisson();
  p->dim = dim;
  p->dtype = dtype;
  p->nz = dim == 3 ? nz : 1;
  p->ny = ny;
  p->nx = nx;
  p->gs = gs;
  p->rank = rank;
  p->nranks = nranks;
  p->backend = backend;
  p->x_range = x_range;
  p->y_range = x_range * ((double)ny / (double)nx);
  p->z_range = x_range * ((double)p->nz / (double)nx);
  p->dx = dtype == SB200_F32 ? (double)(float)(x_range / nx) : x_range / nx;
  p->backend_state = nullptr;
  int e = backend == 1 ? sb_poisson_fft_create(p, stream) : sb_poisson_cufft_create(p, stream);
  if (e) {
    if (backend == 1) sb_poisson_fft_destroy(p); else sb_poisson_cufft_destroy(p);
    delete p;
    return e;
  }
  *out = p;
  return 0;
}

extern "C" int sb200_poisson_destroy(sb200_poisson_t* p) {
  if (!p) return 0;
  if (p->backend == 1) sb_poisson_fft_destroy(p); else sb_poisson_cufft_destroy(p);
  delete p;
  return 0;
}

extern "C" int sb200_poisson_solve(sb200_poisson_t* p, void* solution, const void* rhs, int ncomp,
                                   void* stream) {
  SB_REQUIRE(p && solution && rhs, "poisson_solve: null argument");
  SB_REQUIRE(p->nranks == 1, "poisson_solve: single-rank entry point called on a distributed handle");
  return p->backend == 1 ? sb_poisson_fft_solve(p, solution, rhs, ncomp, stream)
                         : sb_poisson_cufft_solve(p, solution, rhs, ncomp, stream);
}

extern "C" int64_t sb200_poisson_workspace_bytes(const sb200_poisson_t* p) {
  if (!p) return 0;
  return p->backend == 1 ? sb_poisson_fft_bytes(p) : sb_poisson_cufft_bytes(p);
}
