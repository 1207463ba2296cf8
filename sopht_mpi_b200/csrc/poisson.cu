// Poisson handle management, backend dispatch and the Green's function kernel.
#include "poisson.h"

#include <vector>

template <typename T>
struct GreensFillOp {
  T* dst;
  SbGreens<T> g;
  long long n2y, n2x;
  SB_D void operator()(long long i) const {
    const long long x = i % n2x;
    const long long r = i / n2x;
    dst[i] = g(r / n2y, r % n2y, x);
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
int sb_poisson_make_greens(const sb200_poisson* p, SbGreens<T>* op, void** lines_dev, void* stream) {
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
  cudaStreamSynchronize((cudaStream_t)stream);
#else
  d = (T*)malloc(h.size() * sizeof(T));
  memcpy(d, h.data(), h.size() * sizeof(T));
#endif
  *lines_dev = d;
  op->xl = d;
  op->yl = d + lx.size();
  op->zl = d + lx.size() + ly.size();
  op->dim = p->dim;
  op->two_xr = (T)(2 * p->x_range);
  op->two_yr = (T)(2 * p->y_range);
  op->two_zr = (T)(2 * p->z_range);
  op->four_pi = (T)(4 * M_PI);
  op->two_pi = (T)(2 * M_PI);
  if (p->dim == 3)
    op->g0 = (T)(1.0 / (4 * M_PI * p->dx));
  else
    op->g0 = (T)(-(2 * std::log(p->dx / std::sqrt(M_PI)) - 1) / (4 * M_PI));
  return 0;
}
template int sb_poisson_make_greens<float>(const sb200_poisson*, SbGreens<float>*, void**, void*);
template int sb_poisson_make_greens<double>(const sb200_poisson*, SbGreens<double>*, void**, void*);

void sb_poisson_free_greens_lines(void* d) {
#ifndef SB200_EMU
  cudaFree(d);
#else
  free(d);
#endif
}

template <typename T>
static int fill_greens_t(const sb200_poisson* p, void* dst, void* stream) {
  const int n2z = p->dim == 3 ? 2 * p->nz : 1, n2y = 2 * p->ny, n2x = 2 * p->nx;
  GreensFillOp<T> op;
  void* lines = nullptr;
  int e = sb_poisson_make_greens<T>(p, &op.g, &lines, stream);
  if (e) return e;
  op.dst = (T*)dst;
  op.n2y = n2y;
  op.n2x = n2x;
  e = sb_launch_flat((long long)n2z * n2y * n2x, op, stream, "greens");
#ifndef SB200_EMU
  cudaStreamSynchronize((cudaStream_t)stream);
#endif
  sb_poisson_free_greens_lines(lines);
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
