// Shared device-side geometry helpers for libsophtb200.
#pragma once
#include "sb200_rt.h"
#include "../../include/sopht_b200.h"

// error plumbing (thread-local message, returned through sb200_last_error)
void sb_set_error(const char* fmt, ...);
#define SB_CHECK_LAUNCH(what)                                              \
  do {                                                                     \
    int e__ = sb_last_launch_error();                                      \
    if (e__ != 0) {                                                        \
      sb_set_error("%s: %s", what, sb_error_string(e__));                  \
      return -2;                                                           \
    }                                                                      \
  } while (0)
#define SB_REQUIRE(cond, msg) \
  do {                        \
    if (!(cond)) {            \
      sb_set_error("%s", msg); \
      return -1;              \
    }                         \
  } while (0)

// Device view of the local padded grid.
struct SbGeom {
  int dim, gs;
  int mz, my, mx;        // padded sizes (2D: mz = 1)
  int phys[6];           // z_prev,z_next,y_prev,y_next,x_prev,x_next
  long long plane, vol;  // my*mx, mz*my*mx

  SB_HD long long idx(int z, int y, int x) const { return ((long long)z * my + y) * mx + x; }

  // Is (z,y,x) in the union of the seven regions the reference wrappers write
  // for a stencil of support ks?  (interior call + 6 slab calls, SURVEY App. B;
  // e.g. reference diffusion_flux_mpi_3d.py:49-159).  `x_full` reproduces the
  // curl wrapper whose interior call does not slice x (curl_mpi_3d.py:44-48).
  SB_HD bool written(int z, int y, int x, int ks, bool x_full = false) const {
    const bool xs = (x >= gs && x < gs + ks) || (x >= mx - gs - ks && x < mx - gs);
    const bool xi = x >= gs + ks && x < mx - gs - ks;
    const bool ys = (y >= gs && y < gs + ks) || (y >= my - gs - ks && y < my - gs);
    const bool yi = y >= gs + ks && y < my - gs - ks;
    const bool yfull = y >= ks && y < my - ks;
    if (dim == 2) return (xs && yfull) || (xi && (ys || yi));
    const bool zs = (z >= gs && z < gs + ks) || (z >= mz - gs - ks && z < mz - gs);
    const bool zi = z >= gs + ks && z < mz - gs - ks;
    const bool zfull = z >= ks && z < mz - ks;
    bool w = (xs && yfull && zfull) || (xi && ys && zfull) || (xi && yi && (zs || zi));
    if (x_full) w = w || (zi && yi && x >= ks && x < mx - ks);
    return w;
  }
  // physical-boundary ring of width gs+1 that the wrappers zero afterwards
  // (reference diffusion_flux_mpi_3d.py:161-192)
  SB_HD bool in_ring(int z, int y, int x, int width = 1) const {
    const int w = gs + width;
    bool r = (phys[4] && x < w) || (phys[5] && x >= mx - w) || (phys[2] && y < w) ||
             (phys[3] && y >= my - w);
    if (dim == 3) r = r || (phys[0] && z < w) || (phys[1] && z >= mz - w);
    return r;
  }
  // deep interior for support-1 wrappers: written by the interior call and outside every ring
  SB_HD bool deep(int z, int y, int x) const {
    bool r = (unsigned)(x - gs - 1) < (unsigned)(mx - 2 * gs - 2) &&
             (unsigned)(y - gs - 1) < (unsigned)(my - 2 * gs - 2);
    if (dim == 3) r = r && (unsigned)(z - gs - 1) < (unsigned)(mz - 2 * gs - 2);
    return r;
  }
  SB_HD bool interior(int z, int y, int x) const {
    bool r = x >= gs && x < mx - gs && y >= gs && y < my - gs;
    if (dim == 3) r = r && z >= gs && z < mz - gs;
    return r;
  }
};

static inline int sb_make_geom(const sb200_grid_t* g, SbGeom* o) {
  if (!g || (g->dim != 2 && g->dim != 3) || g->gs < 0) return -1;
  o->dim = g->dim;
  o->gs = g->gs;
  o->mz = g->dim == 3 ? g->n[0] + 2 * g->gs : 1;
  o->my = g->n[1] + 2 * g->gs;
  o->mx = g->n[2] + 2 * g->gs;
  for (int i = 0; i < 6; ++i) o->phys[i] = g->phys[i];
  o->plane = (long long)o->my * o->mx;
  o->vol = o->plane * o->mz;
  return 0;
}

// one thread per cell; threads run over the flattened (y,x) plane so that blocks stay full
// whatever the row length, blockIdx.y walks the planes
template <typename Op>
__global__ void __launch_bounds__(256) sb_cell_kernel(SbGeom g, Op op) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int z = blockIdx.y;
  if (idx < g.plane) {
    const int y = (int)(idx / g.mx);
    const int x = (int)(idx - (long long)y * g.mx);
    op(g, z, y, x);
  }
}

template <typename Op>
static inline int sb_launch_cells(const SbGeom& g, const Op& op, void* stream, const char* what) {
  dim3 block(256);
  dim3 grid((unsigned)((g.plane + 255) / 256), (unsigned)g.mz);
  SB_LAUNCH(sb_cell_kernel<Op>, grid, block, 0, stream, g, op);
  SB_CHECK_LAUNCH(what);
  return 0;
}

// a list of up to 6 boxes [lo,hi) of cells; one launch, blockIdx.y selects the box
struct SbBoxes {
  int n;
  int lo[6][3], hi[6][3];  // (z,y,x)
};
template <typename Op>
__global__ void __launch_bounds__(256) sb_box_kernel(SbGeom g, SbBoxes b, Op op) {
  const int k = blockIdx.y;
  const long long nz = b.hi[k][0] - b.lo[k][0], ny = b.hi[k][1] - b.lo[k][1], nx = b.hi[k][2] - b.lo[k][2];
  const long long count = nz * ny * nx;
  if (count < (1LL << 31)) {
    // 32-bit index arithmetic (64-bit division costs ~5x as many instructions; the boxes are face slabs
    // of a few planes, far below 2^31 cells)
    const unsigned n = (unsigned)count, unx = (unsigned)nx, uny = (unsigned)ny;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const unsigned r = i / unx, x = i - r * unx, z = r / uny, y = r - z * uny;
      op(g, b.lo[k][0] + (int)z, b.lo[k][1] + (int)y, b.lo[k][2] + (int)x);
    }
    return;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % nx);
    const long long r = i / nx;
    op(g, b.lo[k][0] + (int)(r / ny), b.lo[k][1] + (int)(r % ny), b.lo[k][2] + x);
  }
}
template <typename Op>
static inline int sb_launch_boxes(const SbGeom& g, const SbBoxes& b, const Op& op, void* stream,
                                  const char* what) {
  long long most = 0;
  for (int k = 0; k < b.n; ++k) {
    long long c = 1;
    for (int d = 0; d < 3; ++d) c *= (b.hi[k][d] > b.lo[k][d] ? b.hi[k][d] - b.lo[k][d] : 0);
    most = c > most ? c : most;
  }
  if (b.n == 0 || most == 0) return 0;
  long long blocks = (most + 255) / 256;
  if (blocks > 2048) blocks = 2048;
  SB_LAUNCH(sb_box_kernel<Op>, dim3((unsigned)blocks, (unsigned)b.n), dim3(256), 0, stream, g, b, op);
  SB_CHECK_LAUNCH(what);
  return 0;
}

// flat kernel over `count` elements
template <typename Op>
__global__ void __launch_bounds__(256) sb_flat_kernel(long long count, Op op) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < count; i += stride) op(i);
}

template <typename Op>
static inline int sb_launch_flat(long long count, const Op& op, void* stream, const char* what) {
  if (count <= 0) return 0;
  long long blocks = (count + 255) / 256;
  const long long cap = 148LL * 32;  // persistent-ish grid: multiple of the SM count
  if (blocks > cap) blocks = cap;
  SB_LAUNCH(sb_flat_kernel<Op>, dim3((unsigned)blocks), dim3(256), 0, stream, count, op);
  SB_CHECK_LAUNCH(what);
  return 0;
}

#define SB_DISPATCH_DTYPE(dtype, CALL)                 \
  do {                                                 \
    if ((dtype) == SB200_F32) {                        \
      using T = float;                                 \
      CALL;                                            \
    } else if ((dtype) == SB200_F64) {                 \
      using T = double;                                \
      CALL;                                            \
    } else {                                           \
      sb_set_error("unsupported dtype %d", (int)(dtype)); \
      return -1;                                       \
    }                                                  \
  } while (0)
