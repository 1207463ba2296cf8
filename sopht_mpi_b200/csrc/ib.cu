// Immersed-boundary kernels: one warp per Lagrangian point, the (2w)^dim
// support cells are spread over the lanes, shuffle-reduced.
// Follows reference numeric/immersed_boundary_ops/EulerianLagrangianGridCommunicatorMPI3D.py
// (:116-178 support/nearest index, :443-589 weights, :181-326 E->L, :329-427 L->E)
// and VirtualBoundaryForcingMPI.py:278-331.
#include "sb200_common.h"
#include <type_traits>

// numpy / numba floor division for floats (npy_divmod): exact fmod based, so
// nearest indices are bit-identical with the reference.
template <typename C>
SB_D C sb_floor_divide(C a, C b) {
  C mod = fmod(a, b);
  C div = (a - mod) / b;
  if (mod != C(0)) {
    if ((b < C(0)) != (mod < C(0))) {
      mod += b;
      div -= C(1);
    }
  }
  C floordiv;
  if (div != C(0)) {
    floordiv = floor(div);
    if (div - floordiv > C(0.5)) floordiv += C(1);
  } else {
    floordiv = C(0);
  }
  return floordiv;
}

template <typename TL>
struct SbIbP {
  int dim, width, kernel_type;
  long long sub_shift[3];  // substart_xyz - gs
  double dx, shift;
  TL weight_prefac;  // real_t((0.25/dx)**dim) or (0.125/dx)**dim
  TL half_pi;        // real_t(0.5*pi)
};

// 1D delta-kernel factor for scaled distance r = support/dx
template <typename TL>
SB_D TL sb_delta_1d(const SbIbP<TL>& P, TL s_over_dx) {
  if (P.kernel_type == 0) return TL(1) + cos(P.half_pi * s_over_dx);
  const TL r = fabs(s_over_dx);
  TL v = 0;
  if (r < TL(1))
    v = TL(3) - TL(2) * r + sqrt(fabs(TL(1) + TL(4) * r - TL(4) * r * r));
  else if (r < TL(2))
    v = TL(5) - TL(2) * r - sqrt(fabs(TL(-7) + TL(12) * r - TL(4) * r * r));
  return v;
}

// nearest index (local padded frame) and scaled support offset of cell k along axis d
template <typename TL, typename TC>
SB_D long long sb_nearest(const SbIbP<TL>& P, TL pos, int d) {
  const TC q = sb_floor_divide<TC>((TC)pos - (TC)P.shift, (TC)P.dx);
  return (long long)q - P.sub_shift[d];
}
template <typename TL>
SB_D TL sb_support_scaled(const SbIbP<TL>& P, long long nearest, int off, TL pos, int d) {
  // (idx + off + sub_shift) * dx + shift - pos  evaluated in double, stored as TL, then /= dx
  const double s = (double)(nearest + off + P.sub_shift[d]) * P.dx + P.shift - (double)pos;
  return (TL)s / (TL)P.dx;
}

// Eight lanes per Lagrangian point (four points per warp): lane `sub` of a group owns the window cells
//   3D: kz = sub >> 1, ky in {2 (sub & 1), 2 (sub & 1) + 1}, kx = 0..3      (8 of the 64 cells)
//   2D: ky = sub >> 1, kx in {2 (sub & 1), 2 (sub & 1) + 1}                  (2 of the 16 cells)
// The twelve 1D delta factors of a point are evaluated once (lane sub: factor `sub`, lanes 0..3 also
// factor 8 + sub) and handed round the group with shuffles; the per-cell weight is the product
// prefac * f_x * f_y * f_z in the reference's order, so the weights are unchanged bit for bit.
template <typename TL>
struct SbIbCells {
  int kx[8], ky[8], kz[8], n;
};
template <typename TE, typename TL>
SB_D void sb_group_weights(const SbIbP<TL>& P, const long long (&near)[3], const TL (&p)[3], unsigned lane,
                           TL (&w)[8], int (&ck)[8][3], int& ncell) {
  const unsigned sub = lane & 7u, base = lane & ~7u;
  // factor `sub` (axes x: 0..3, y: 4..7) and, on lanes 0..3, factor 8 + sub (axis z)
  TL fa, fb = TL(0);
  {
    const int d = (int)(sub >> 2), k = (int)(sub & 3);
    fa = sb_delta_1d(P, sb_support_scaled(P, d == 0 ? near[0] : near[1], k - P.width + 1, d == 0 ? p[0] : p[1], d));
    if (P.dim == 3 && sub < 4) fb = sb_delta_1d(P, sb_support_scaled(P, near[2], (int)sub - P.width + 1, p[2], 2));
  }
  TL fx[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) fx[k] = __shfl_sync(0xffffffffu, fa, base + k);
  if (P.dim == 3) {
    const int kz = (int)(sub >> 1), ky0 = 2 * (int)(sub & 1);
    const TL fy0 = __shfl_sync(0xffffffffu, fa, base + 4 + ky0), fy1 = __shfl_sync(0xffffffffu, fa, base + 5 + ky0);
    const TL fz = __shfl_sync(0xffffffffu, fb, base + kz);
    ncell = 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int kx = i & 3, ky = ky0 + (i >> 2);
      TL v = P.weight_prefac * fx[kx];
      v = v * (i < 4 ? fy0 : fy1);
      v = v * fz;
      // the reference's Peskin kernel hands back weights rounded to real_t even when the Lagrangian
      // arrays are wider (tests/golden/ib_*_f32_f64.npz: every w_pes value is a float32 number)
      if (P.kernel_type == 1) v = (TL)(TE)v;
      w[i] = v;
      ck[i][0] = kx;
      ck[i][1] = ky;
      ck[i][2] = kz;
    }
  } else {
    const int ky = (int)(sub >> 1), kx0 = 2 * (int)(sub & 1);
    const TL fy = __shfl_sync(0xffffffffu, fa, base + 4 + ky);
    ncell = 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int kx = kx0 + (i & 1);
      TL v = P.weight_prefac * ((i & 1) ? (kx0 ? fx[3] : fx[1]) : (kx0 ? fx[2] : fx[0]));
      v = v * fy;
      if (P.kernel_type == 1) v = (TL)(TE)v;
      w[i] = v;
      ck[i][0] = kx;
      ck[i][1] = ky;
      ck[i][2] = 0;
    }
  }
}

template <typename TE, typename TL, typename TC>
__global__ void __launch_bounds__(128)
    sb_ib_interact_kernel(SbGeom g, SbIbP<TL> P, long long n, int ncomp, const TE* eul, const TL* pos,
                          const TL* vel, const TL* dpos, long long* nearest_out, TL* weights_out,
                          TL* flow_vel, TL* dvel, TL* force, double dx_pow_dim, TL kcoef, TL ccoef,
                          const int* __restrict__ owner, int my_rank) {
  const unsigned lane = threadIdx.x & 31, sub = lane & 7u;
  const long long pt0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 4 + (lane >> 3);
  const bool live = pt0 < n;
  const long long pt = live ? pt0 : n - 1;  // (idle groups repeat the last point and store nothing)
  const bool mine = !(owner && owner[pt] != my_rank);
  const int dim = P.dim, kw = 2 * P.width;
  long long near[3] = {0, 0, 0};
  TL p[3] = {0, 0, 0};
  for (int d = 0; d < dim; ++d) {
    p[d] = pos[d * n + pt];
    near[d] = sb_nearest<TL, TC>(P, p[d], d);
  }
  TL w[8];
  int ck[8][3], ncell;
  sb_group_weights<TE, TL>(P, near, p, lane, w, ck, ncell);
  double acc[3] = {0.0, 0.0, 0.0};
  const int off0 = -P.width + 1;
  if (live && mine) {
    if (nearest_out && sub == 0)
      for (int d = 0; d < dim; ++d) nearest_out[d * n + pt] = near[d];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i >= ncell) break;
      const int cell = (ck[i][2] * kw + ck[i][1]) * kw + ck[i][0];
      if (weights_out) weights_out[(long long)cell * n + pt] = w[i];
      const long long x = near[0] + ck[i][0] + off0, y = near[1] + ck[i][1] + off0;
      const long long z = dim == 3 ? near[2] + ck[i][2] + off0 : 0;
      if (x >= 0 && x < g.mx && y >= 0 && y < g.my && z >= 0 && z < g.mz) {
        const long long idx = (z * g.my + y) * g.mx + x;
        for (int c = 0; c < ncomp; ++c) acc[c] += (double)eul[idx + c * g.vol] * (double)w[i];
      }
    }
  }
  for (int c = 0; c < ncomp; ++c)
    for (int o = 4; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
  if (live && sub == 0) {
    for (int c = 0; c < ncomp; ++c) {
      if (!mine) {
        // another rank's point: zeros, so that a SUM all-reduce over the ranks assembles the global arrays
        flow_vel[c * n + pt] = TL(0);
        if (force) dvel[c * n + pt] = force[c * n + pt] = TL(0);
        continue;
      }
      const TL u = (TL)(acc[c] * dx_pow_dim);
      flow_vel[c * n + pt] = u;
      if (force) {
        const TL dv = u - vel[c * n + pt];
        dvel[c * n + pt] = dv;
        force[c * n + pt] = kcoef * dpos[c * n + pt] + ccoef * dv;
      }
    }
  }
}

template <typename TE, typename TL, typename TC>
__global__ void __launch_bounds__(128)
    sb_ib_spread_kernel(SbGeom g, SbIbP<TL> P, long long n, TE* eul, const TL* lag, const TL* pos,
                        const int* __restrict__ owner, int my_rank) {
  const unsigned lane = threadIdx.x & 31;
  const long long pt0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 4 + (lane >> 3);
  const bool live = pt0 < n;
  const long long pt = live ? pt0 : n - 1;
  const bool mine = !(owner && owner[pt] != my_rank);
  const int dim = P.dim;
  long long near[3] = {0, 0, 0};
  TL p[3] = {0, 0, 0}, f[3] = {0, 0, 0};
  for (int d = 0; d < dim; ++d) {
    p[d] = pos[d * n + pt];
    f[d] = lag[d * n + pt];
    near[d] = sb_nearest<TL, TC>(P, p[d], d);
  }
  TL w[8];
  int ck[8][3], ncell;
  sb_group_weights<TE, TL>(P, near, p, lane, w, ck, ncell);
  if (!(live && mine)) return;
  const int off0 = -P.width + 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i >= ncell) break;
    const long long x = near[0] + ck[i][0] + off0, y = near[1] + ck[i][1] + off0;
    const long long z = dim == 3 ? near[2] + ck[i][2] + off0 : 0;
    if (x >= 0 && x < g.mx && y >= 0 && y < g.my && z >= 0 && z < g.mz) {
      const long long idx = (z * g.my + y) * g.mx + x;
      for (int c = 0; c < dim; ++c) atomicAdd(&eul[idx + c * g.vol], (TE)(f[c] * w[i]));
    }
  }
}

template <typename TL>
static int sb_make_ib(const sb200_grid_t* gr, const sb200_ib_params_t* p, SbIbP<TL>* o) {
  o->dim = gr->dim;
  o->width = p->width;
  o->kernel_type = p->kernel_type;
  for (int d = 0; d < 3; ++d) o->sub_shift[d] = (long long)p->substart_xyz[d] - gr->gs;
  o->dx = p->dx;
  o->shift = p->coord_shift;
  // constants rounded through real_t exactly as the reference does
  // (...MPI3D.py:464-477: real_t((0.25/dx)**dim), real_t(0.5*pi))
  const double base = (p->kernel_type == 0 ? 0.25 : 0.125) / p->dx;
  double pref = base;
  for (int d = 1; d < gr->dim; ++d) pref *= base;
  if (gr->dtype == SB200_F32) {
    o->weight_prefac = (TL)(float)pref;
    o->half_pi = (TL)(float)(0.5 * M_PI);
  } else {
    o->weight_prefac = (TL)pref;
    o->half_pi = (TL)(0.5 * M_PI);
  }
  return 0;
}

template <typename TE, typename TL>
static int ib_interact_t(const sb200_grid_t* gr, const sb200_ib_params_t* p, long long n, int ncomp,
                         const void* eul, const void* pos, const void* vel, const void* dpos,
                         void* nearest, void* weights, void* flow_vel, void* dvel, void* force,
                         void* stream, const int* owner = nullptr, int my_rank = 0) {
  // index arithmetic in the promoted type of (lag dtype, real_t)
  using TC = typename std::conditional<(sizeof(TE) > sizeof(TL)), TE, TL>::type;
  SbGeom g;
  SB_REQUIRE(sb_make_geom(gr, &g) == 0, "bad grid");
  SbIbP<TL> P;
  sb_make_ib<TL>(gr, p, &P);
  if (n <= 0) return 0;
  double dxp = 1.0;
  // reference: dx**grid_dim with dx a real_t scalar
  if (gr->dtype == SB200_F32) {
    float d = (float)p->dx, a = d;
    for (int k = 1; k < gr->dim; ++k) a *= d;
    dxp = a;
  } else {
    for (int k = 0; k < gr->dim; ++k) dxp *= p->dx;
  }
  const int warps = 4, per_block = 4 * warps;  // four points per warp
  dim3 block(32 * warps), grid((unsigned)((n + per_block - 1) / per_block));
  SB_LAUNCH_COOP((sb_ib_interact_kernel<TE, TL, TC>), grid, block, 0, stream, g, P, n, ncomp,
                 (const TE*)eul, (const TL*)pos, (const TL*)vel, (const TL*)dpos, (long long*)nearest,
                 (TL*)weights, (TL*)flow_vel, (TL*)dvel, (TL*)force, dxp, (TL)p->stiffness,
                 (TL)p->damping, owner, my_rank);
  SB_CHECK_LAUNCH("ib_interact");
  return 0;
}

template <typename TE, typename TL>
static int ib_spread_t(const sb200_grid_t* gr, const sb200_ib_params_t* p, long long n, void* eul,
                       const void* lag, const void* pos, void* stream, const int* owner = nullptr,
                       int my_rank = 0) {
  using TC = typename std::conditional<(sizeof(TE) > sizeof(TL)), TE, TL>::type;
  SbGeom g;
  SB_REQUIRE(sb_make_geom(gr, &g) == 0, "bad grid");
  SbIbP<TL> P;
  sb_make_ib<TL>(gr, p, &P);
  if (n <= 0) return 0;
  const int warps = 4, per_block = 4 * warps;
  dim3 block(32 * warps), grid((unsigned)((n + per_block - 1) / per_block));
  SB_LAUNCH_COOP((sb_ib_spread_kernel<TE, TL, TC>), grid, block, 0, stream, g, P, n, (TE*)eul,
                 (const TL*)lag, (const TL*)pos, owner, my_rank);
  SB_CHECK_LAUNCH("ib_spread");
  return 0;
}

#define SB_DISPATCH_2(edt, ldt, CALL)                                   \
  do {                                                                  \
    if ((edt) == SB200_F32 && (ldt) == SB200_F32) {                     \
      using TE = float; using TL = float; CALL;                         \
    } else if ((edt) == SB200_F32 && (ldt) == SB200_F64) {              \
      using TE = float; using TL = double; CALL;                        \
    } else if ((edt) == SB200_F64 && (ldt) == SB200_F64) {              \
      using TE = double; using TL = double; CALL;                       \
    } else if ((edt) == SB200_F64 && (ldt) == SB200_F32) {              \
      using TE = double; using TL = float; CALL;                        \
    } else {                                                            \
      sb_set_error("unsupported dtype combination");                    \
      return -1;                                                        \
    }                                                                   \
  } while (0)

extern "C" int sb200_ib_interact_lag(const sb200_grid_t* g, const sb200_ib_params_t* p, int64_t n,
                                     const void* eul_velocity, const void* lag_position,
                                     const void* lag_velocity, const void* position_mismatch,
                                     void* nearest, void* weights, void* flow_velocity,
                                     void* velocity_mismatch, void* forcing, void* stream) {
  SB_REQUIRE(g && p, "ib: null params");
  SB_REQUIRE(p->width == 2, "Interpolation kernel inconsistent with interpolation kernel width!");
  SB_REQUIRE(g->gs >= p->width, "ghost size needs to be >= interp kernel width");
  SB_DISPATCH_2(g->dtype, p->lag_dtype,
                return (ib_interact_t<TE, TL>(g, p, n, g->dim, eul_velocity, lag_position, lag_velocity,
                                              position_mismatch, nearest, weights, flow_velocity,
                                              velocity_mismatch, forcing, stream)));
}

extern "C" int sb200_ib_interpolate(const sb200_grid_t* g, const sb200_ib_params_t* p, int64_t n, int ncomp,
                                    const void* eul_field, const void* lag_position, void* lag_field,
                                    void* stream) {
  SB_REQUIRE(g && p, "ib: null params");
  SB_REQUIRE(p->width == 2, "Interpolation kernel inconsistent with interpolation kernel width!");
  SB_REQUIRE(ncomp == 1 || ncomp == g->dim, "invalid number of components for interpolation!");
  SB_DISPATCH_2(g->dtype, p->lag_dtype,
                return (ib_interact_t<TE, TL>(g, p, n, ncomp, eul_field, lag_position, nullptr, nullptr,
                                              nullptr, nullptr, lag_field, nullptr, nullptr, stream)));
}

extern "C" int sb200_ib_spread(const sb200_grid_t* g, const sb200_ib_params_t* p, int64_t n,
                               void* eul_forcing, const void* lag_forcing, const void* lag_position,
                               void* stream) {
  SB_REQUIRE(g && p, "ib: null params");
  SB_REQUIRE(p->width == 2, "Interpolation kernel inconsistent with interpolation kernel width!");
  SB_DISPATCH_2(g->dtype, p->lag_dtype,
                return (ib_spread_t<TE, TL>(g, p, n, eul_forcing, lag_forcing, lag_position, stream)));
}

// ---------------------------------------------- Lagrangian rank ownership (L1) on the device
// rank_address[i] = rank_map[cz, cy, cx] with the block coordinate of every array axis
//   ((pos - shift) / (dx * n_local_axis)).astype(int32)      (truncation toward zero)
// evaluated like numpy evaluates the reference expression (utils/mpi_utils_3d.py:1357-1384,
// mpi_utils_2d.py:611-637): the subtraction in the dtype of the positions, the division in double
// (`eul_subblock_dx` is a float64 array).  `flag` (device int, may be NULL) is set to 1 when a point
// lies beyond the topology (the reference aborts; the caller raises).
template <typename TL>
struct RankAddressOp {
  const TL* pos;
  int* out;
  int* flag;
  long long n;
  int dim;
  int topo[3];      // array order (z,y,x); 2D: topo[0] = 1
  double sub_dx[3];
  TL shift;
  SB_D void operator()(long long i) const {
    int rank = 0;
    bool bad = false;
    for (int ax = 3 - dim; ax < 3; ++ax) {
      const TL d = pos[(long long)(2 - ax) * n + i] - shift;  // positions are x,y,z: component 2 - ax
      int c = (int)((double)d / sub_dx[ax]);
      if (c >= topo[ax]) bad = true;
      if (c < 0) c += topo[ax];  // numpy's negative-index wrap of rank_map[...]
      if (c < 0 || c >= topo[ax]) {
        bad = true;
        c = 0;
      }
      rank = rank * topo[ax] + c;
    }
    out[i] = rank;
    if (bad && flag) *flag = 1;
  }
};

template <typename T>
static int ib_rank_address_t(int dim, long long n, const void* pos, double shift, const double* sub_dx,
                             const int32_t* topo, void* out, void* flag, void* stream) {
  RankAddressOp<T> op;
  op.pos = (const T*)pos;
  op.out = (int*)out;
  op.flag = (int*)flag;
  op.n = n;
  op.dim = dim;
  for (int a = 0; a < 3; ++a) {
    op.topo[a] = topo[a];
    op.sub_dx[a] = sub_dx[a];
  }
  op.shift = (T)shift;
  return sb_launch_flat(n, op, stream, "ib_rank_address");
}
extern "C" int sb200_ib_rank_address(int lag_dtype, int dim, int64_t n, const void* lag_position,
                                     double coord_shift, const double* sub_dx_zyx, const int32_t* topo_zyx,
                                     void* rank_address, void* out_of_domain_flag, void* stream) {
  SB_REQUIRE(dim == 2 || dim == 3, "ib_rank_address: dim must be 2 or 3");
  SB_REQUIRE(lag_position && rank_address && sub_dx_zyx && topo_zyx, "ib_rank_address: null pointer");
  if (n <= 0) return 0;
  SB_DISPATCH_DTYPE(lag_dtype, return ib_rank_address_t<T>(dim, n, lag_position, coord_shift, sub_dx_zyx, topo_zyx,
                                                           rank_address, out_of_domain_flag, stream));
}

// the same kernels restricted to the points this rank owns (`rank_address[i] == my_rank`); the
// interaction writes zeros for the others (VirtualBoundaryForcingMPI keeps replicated global arrays
// and assembles them with one SUM all-reduce instead of the reference's master scatter / gather,
// utils/mpi_utils_3d.py:1386-1459)
extern "C" int sb200_ib_interact_owned(const sb200_grid_t* g, const sb200_ib_params_t* p, int64_t n,
                                       const void* eul_velocity, const void* lag_position,
                                       const void* lag_velocity, const void* position_mismatch, void* nearest,
                                       void* weights, void* flow_velocity, void* velocity_mismatch,
                                       void* forcing, const void* rank_address, int my_rank, void* stream) {
  SB_REQUIRE(g && p, "ib: null params");
  SB_REQUIRE(p->width == 2, "Interpolation kernel inconsistent with interpolation kernel width!");
  SB_REQUIRE(g->gs >= p->width, "ghost size needs to be >= interp kernel width");
  SB_DISPATCH_2(g->dtype, p->lag_dtype,
                return (ib_interact_t<TE, TL>(g, p, n, g->dim, eul_velocity, lag_position, lag_velocity,
                                              position_mismatch, nearest, weights, flow_velocity,
                                              velocity_mismatch, forcing, stream, (const int*)rank_address,
                                              my_rank)));
}
extern "C" int sb200_ib_spread_owned(const sb200_grid_t* g, const sb200_ib_params_t* p, int64_t n,
                                     void* eul_forcing, const void* lag_forcing, const void* lag_position,
                                     const void* rank_address, int my_rank, void* stream) {
  SB_REQUIRE(g && p, "ib: null params");
  SB_REQUIRE(p->width == 2, "Interpolation kernel inconsistent with interpolation kernel width!");
  SB_DISPATCH_2(g->dtype, p->lag_dtype,
                return (ib_spread_t<TE, TL>(g, p, n, eul_forcing, lag_forcing, lag_position, stream,
                                            (const int*)rank_address, my_rank)));
}

// ------------------------------------------------------------ ghost cells --
template <typename T>
struct ClearGhostOp {
  T* f;
  int ncomp;
  SB_D void operator()(const SbGeom& g, int z, int y, int x) const {
    if (g.interior(z, y, x)) return;
    const long long i = g.idx(z, y, x);
    for (int c = 0; c < ncomp; ++c) f[i + c * g.vol] = 0;
  }
};
extern "C" int sb200_clear_ghost_cells(const sb200_grid_t* gr, void* field, int ncomp, void* stream) {
  SbGeom g;
  SB_REQUIRE(sb_make_geom(gr, &g) == 0, "bad grid");
  if (g.gs == 0) return 0;
  // only the six ghost slabs are visited (disjoint boxes: x slabs, y slabs between them, z slabs in
  // the middle), not every cell of the field
  SbBoxes b;
  b.n = 0;
  auto add = [&](int z0, int z1, int y0, int y1, int x0, int x1) {
    const int lo[3] = {z0, y0, x0}, hi[3] = {z1, y1, x1};
    for (int d = 0; d < 3; ++d) {
      b.lo[b.n][d] = lo[d];
      b.hi[b.n][d] = hi[d];
    }
    ++b.n;
  };
  const int gs = g.gs;
  add(0, g.mz, 0, g.my, 0, gs);
  add(0, g.mz, 0, g.my, g.mx - gs, g.mx);
  add(0, g.mz, 0, gs, gs, g.mx - gs);
  add(0, g.mz, g.my - gs, g.my, gs, g.mx - gs);
  if (g.dim == 3) {
    add(0, gs, gs, g.my - gs, gs, g.mx - gs);
    add(g.mz - gs, g.mz, gs, g.my - gs, gs, g.mx - gs);
  }
  SB_DISPATCH_DTYPE(gr->dtype,
                    return sb_launch_boxes(g, b, ClearGhostOp<T>{(T*)field, ncomp}, stream, "clear_ghost"));
}

// field[gs:2gs] += from_prev ; field[-2gs:-gs] += from_next   (z slabs of gs padded planes,
// per component; either pointer may be NULL).  Reference ...MPI3D.py:689-760 for z-slab topologies.
template <typename T>
struct GhostSumAddOp {
  T* f;
  const T* prev;
  const T* next;
  long long n, slab;  // slab = gs*plane
  long long vol, lo_off, hi_off;
  SB_D void operator()(long long i) const {
    const long long c = i / slab, r = i - c * slab;
    if (prev) f[c * vol + lo_off + r] += prev[i];
    if (next) f[c * vol + hi_off + r] += next[i];
  }
};
extern "C" int sb200_ghost_sum_add_z(const sb200_grid_t* gr, void* field, int ncomp, const void* from_prev,
                                     const void* from_next, void* stream) {
  SbGeom g;
  SB_REQUIRE(sb_make_geom(gr, &g) == 0, "bad grid");
  // slab axis: z in 3D, y in 2D (leading array axis)
  const long long line = g.dim == 3 ? g.plane : g.mx;
  const long long lead = g.dim == 3 ? g.mz : g.my;
  const long long slab = (long long)g.gs * line;
  const long long count = slab * ncomp;
  SB_DISPATCH_DTYPE(gr->dtype, return sb_launch_flat(
                                   count,
                                   GhostSumAddOp<T>{(T*)field, (const T*)from_prev, (const T*)from_next,
                                                    count, slab, g.vol, (long long)g.gs * line,
                                                    (lead - 2 * g.gs) * line},
                                   stream, "ghost_sum_add"));
}

template <typename T>
struct MismatchOp {
  T* dx;
  const T* dv;
  T dt;
  SB_D void operator()(long long i) const { dx[i] = dx[i] + dt * dv[i]; }
};
extern "C" int sb200_ib_update_position_mismatch(int lag_dtype, void* position_mismatch,
                                                 const void* velocity_mismatch, int64_t count, double dt,
                                                 void* stream) {
  SB_DISPATCH_DTYPE(lag_dtype, return sb_launch_flat(count,
                                                     MismatchOp<T>{(T*)position_mismatch,
                                                                   (const T*)velocity_mismatch, (T)dt},
                                                     stream, "ib_mismatch"));
}
