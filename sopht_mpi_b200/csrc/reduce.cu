// Interior reductions (stable-dt / diagnostics) and the fused
// "velocity from stream function" sweep.
#include "sb200_common.h"

// order-preserving map double -> uint64 so that atomicMax works for signed values
SB_D unsigned long long sb_key_from_double(double v) {
  unsigned long long b;
  memcpy(&b, &v, 8);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
SB_D double sb_double_from_key(unsigned long long k) {
  unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  double v;
  memcpy(&v, &b, 8);
  return v;
}

// block-wide max / sum of one double per thread; result valid in thread 0
template <bool IsMax>
SB_D double sb_block_reduce(double v) {
  __shared__ double sh[32];
  const unsigned tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const unsigned nthreads = blockDim.x * blockDim.y * blockDim.z;
  const unsigned lane = tid & 31, warp = tid >> 5;
  for (int o = 16; o > 0; o >>= 1) {
    const double other = __shfl_xor_sync(0xffffffffu, v, o);
    v = IsMax ? (other > v ? other : v) : v + other;
  }
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  if (warp == 0) {
    const unsigned nw = (nthreads + 31) / 32;
    v = lane < nw ? sh[lane] : (IsMax ? -1.0e300 : 0.0);
    for (int o = 16; o > 0; o >>= 1) {
      const double other = __shfl_xor_sync(0xffffffffu, v, o);
      v = IsMax ? (other > v ? other : v) : v + other;
    }
  }
  return v;
}

// mode 0: max of sum_c |f_c| ; 1: signed max over comps ; 2: sum of squares
template <typename T, int MODE, int NCOMP>
__global__ void __launch_bounds__(256) sb_reduce_kernel(SbGeom g, const T* __restrict__ f, int ncomp_rt, void* out,
                                                        int zchunk) {
  // a thread owns one interior (y,x) column and marches over zchunk planes: the (y,x) test is done once,
  // the loads of the unrolled march are independent, one block reduction + one atomic per block
  const int ncomp = NCOMP > 0 ? NCOMP : ncomp_rt;
  const long long pidx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int y = pidx < g.plane ? (int)(pidx / g.mx) : g.my;
  const int x = pidx < g.plane ? (int)(pidx - (long long)y * g.mx) : g.mx;
  const int zlo = g.dim == 3 ? g.gs : 0, zhi = g.dim == 3 ? g.mz - g.gs : 1;  // interior planes
  int zb = blockIdx.y * zchunk, ze = zb + zchunk;
  zb = zb > zlo ? zb : zlo;
  ze = ze < zhi ? ze : zhi;
  double v = MODE == 2 ? 0.0 : -1.0e300;
  if (x >= g.gs && x < g.mx - g.gs && y >= g.gs && y < g.my - g.gs) {
    const T* p = f + g.idx(zb, y, x);
    T acc = sizeof(T) == 4 ? T(-3.402823466e38) : T(-1.7976931348623157e308);  // lowest finite value
#pragma unroll 4
    for (int z = zb; z < ze; ++z, p += g.plane) {
      if (MODE == 0) {
        T s = 0;
#pragma unroll
        for (int c = 0; c < ncomp; ++c) s += fabs(p[c * g.vol]);
        acc = s > acc ? s : acc;
      } else if (MODE == 1) {
#pragma unroll
        for (int c = 0; c < ncomp; ++c) {
          const T t = p[c * g.vol];
          acc = t > acc ? t : acc;
        }
      } else {
#pragma unroll
        for (int c = 0; c < ncomp; ++c) {
          const double t = (double)p[c * g.vol];
          v += t * t;
        }
      }
    }
    if (MODE != 2 && ze > zb) v = (double)acc;
  }
  v = sb_block_reduce<MODE != 2>(v);
  const unsigned tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  if (tid == 0) {
    if (MODE == 2)
      atomicAdd((double*)out, v);
    else
      atomicMax((unsigned long long*)out, sb_key_from_double(v));
  }
}

__global__ void sb_decode_key_kernel(void* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const unsigned long long k = *(unsigned long long*)out;
    *(double*)out = sb_double_from_key(k);
  }
}

template <int MODE>
static int sb_reduce(const sb200_grid_t* gr, const void* field, int ncomp, void* out, void* stream) {
  SbGeom g;
  SB_REQUIRE(sb_make_geom(gr, &g) == 0, "bad grid");
  SB_REQUIRE(out && field, "reduce: null pointer");
  int e = sb_memset_async(out, 0, 8, stream);
  SB_REQUIRE(e == 0, "reduce: memset failed");
  dim3 block(256);
  const int zchunk = g.mz >= 64 ? 16 : (g.mz >= 8 ? 4 : 1);
  dim3 grid((unsigned)((g.plane + 255) / 256), (unsigned)((g.mz + zchunk - 1) / zchunk));
#define SB_LAUNCH_REDUCE(T, NC) \
  SB_LAUNCH_COOP((sb_reduce_kernel<T, MODE, NC>), grid, block, 0, stream, g, (const T*)field, ncomp, out, zchunk)
  if (gr->dtype == SB200_F32) {
    if (ncomp == 3) SB_LAUNCH_REDUCE(float, 3); else if (ncomp == 1) SB_LAUNCH_REDUCE(float, 1); else SB_LAUNCH_REDUCE(float, 0);
  } else {
    if (ncomp == 3) SB_LAUNCH_REDUCE(double, 3); else if (ncomp == 1) SB_LAUNCH_REDUCE(double, 1); else SB_LAUNCH_REDUCE(double, 0);
  }
#undef SB_LAUNCH_REDUCE
  SB_CHECK_LAUNCH("reduce");
  if (MODE != 2) {
    SB_LAUNCH(sb_decode_key_kernel, dim3(1), dim3(32), 0, stream, out);
    SB_CHECK_LAUNCH("reduce_decode");
  }
  return 0;
}

extern "C" int sb200_max_abs_sum(const sb200_grid_t* g, const void* f, int ncomp, void* out, void* stream) {
  return sb_reduce<0>(g, f, ncomp, out, stream);
}
extern "C" int sb200_max(const sb200_grid_t* g, const void* f, int ncomp, void* out, void* stream) {
  return sb_reduce<1>(g, f, ncomp, out, stream);
}
extern "C" int sb200_sum_squares(const sb200_grid_t* g, const void* f, int ncomp, void* out, void* stream) {
  return sb_reduce<2>(g, f, ncomp, out, stream);
}

// --------------------------------------------------------------------------
// u = p*curl(psi) on the wrapper's cells, 0 on the physical ring, + U_inf;
// F = 0; max over interior of sum_c |u_c|  -- one read of psi, one write of u.
// (reference flow_simulators_mpi_3d.py:388-393, 422-424, 429-442)
// Each thread owns one (y,x) column and marches over a chunk of z planes: the (y,x) part of
// the wrapper masks and the addresses are computed once, the z part is uniform per plane.
template <typename T>
__global__ void __launch_bounds__(256)
    sb_velocity_kernel(SbGeom g, T* __restrict__ u, const T* __restrict__ psi, T p, T u0, T u1, T u2,
                       T* __restrict__ forcing, void* max_out, int zchunk) {
  const long long pidx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int zb = blockIdx.y * zchunk;
  const int ze = zb + zchunk < g.mz ? zb + zchunk : g.mz;
  double vmax = -1.0e300;
  if (pidx < g.plane) {
    const int y = (int)(pidx / g.mx);
    const int x = (int)(pidx - (long long)y * g.mx);
    const int gs = g.gs;
    const bool xs = (x == gs) || (x == g.mx - gs - 1);
    const bool xi = x > gs && x < g.mx - gs - 1;
    const bool ys = (y == gs) || (y == g.my - gs - 1);
    const bool yi = y > gs && y < g.my - gs - 1;
    const bool yfull = y >= 1 && y < g.my - 1;
    const bool d3 = g.dim == 3;
    // written = (mA && zfull) || (mB && zin) || (mC && zi)   (2D: mA only)
    const bool mA = d3 ? ((xs && yfull) || (xi && ys)) : ((xs && yfull) || (xi && (ys || yi)));
    const bool mB = d3 && xi && yi;
    const bool mC = d3 && yi && x >= 1 && x < g.mx - 1;  // curl's interior call does not slice x
    const int w = gs + 1;
    const bool ring_yx = (g.phys[4] && x < w) || (g.phys[5] && x >= g.mx - w) || (g.phys[2] && y < w) ||
                         (g.phys[3] && y >= g.my - w);
    const bool int_yx = x >= gs && x < g.mx - gs && y >= gs && y < g.my - gs;
    const long long n = g.vol, sy = g.mx, sz = g.plane;
    long long i = (long long)zb * g.plane + pidx;
    for (int z = zb; z < ze; ++z, i += sz) {
      const bool zfull = !d3 || (z >= 1 && z < g.mz - 1);
      const bool zin = z >= gs && z < g.mz - gs;
      const bool zi = z > gs && z < g.mz - gs - 1;
      const bool ring = ring_yx || (d3 && ((g.phys[0] && z < w) || (g.phys[1] && z >= g.mz - w)));
      const bool wr = (mA && zfull) || (mB && zin) || (mC && zi);
      T c0 = 0, c1 = 0, c2 = 0;
      if (ring || wr) {
        if (!ring) {
          if (d3) {
            const T* fx = psi;
            const T* fy = psi + n;
            const T* fz = psi + 2 * n;
            c0 = p * (fz[i + sy] - fz[i - sy] - fy[i + sz] + fy[i - sz]);
            c1 = p * (fx[i + sz] - fx[i - sz] - fz[i + 1] + fz[i - 1]);
            c2 = p * (fy[i + 1] - fy[i - 1] - fx[i + sy] + fx[i - sy]);
          } else {
            c0 = p * (psi[i + sy] - psi[i - sy]);
            c1 = -p * (psi[i + 1] - psi[i - 1]);
          }
        }
      } else {
        c0 = u[i];
        c1 = u[i + n];
        if (d3) c2 = u[i + 2 * n];
      }
      c0 += u0;
      c1 += u1;
      u[i] = c0;
      u[i + n] = c1;
      T s = fabs(c0) + fabs(c1);
      if (d3) {
        c2 += u2;
        u[i + 2 * n] = c2;
        s += fabs(c2);
      }
      if (forcing) {
        forcing[i] = 0;
        forcing[i + n] = 0;
        if (d3) forcing[i + 2 * n] = 0;
      }
      if (int_yx && (!d3 || zin)) vmax = (double)s > vmax ? (double)s : vmax;
    }
  }
  if (max_out) {
    vmax = sb_block_reduce<true>(vmax);
    if (threadIdx.x == 0) atomicMax((unsigned long long*)max_out, sb_key_from_double(vmax));
  }
}

// Vectorised 3D variant: a thread owns FOUR consecutive x cells of one (y, x) column group and marches
// over z; all neighbour planes come in as 16-byte loads (rows are 16-byte aligned when mx % 4 == 0),
// only the two x-neighbours outside the group are scalar.  Same masks, same arithmetic, same
// reduction as sb_velocity_kernel.
template <typename T>
struct SbVec4;
template <>
struct alignas(16) SbVec4<float> {
  float v[4];
};
template <>
struct alignas(32) SbVec4<double> {
  double v[4];
};

template <typename T>
__global__ void __launch_bounds__(256)
    sb_velocity_vec4_kernel(SbGeom g, T* __restrict__ u, const T* __restrict__ psi, T p, T u0, T u1, T u2,
                            T* __restrict__ forcing, void* max_out, int zchunk) {
  using V = SbVec4<T>;
  const int mx4 = g.mx >> 2;
  const long long gidx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (y, x4) flattened
  const int zb = blockIdx.y * zchunk;
  const int ze = zb + zchunk < g.mz ? zb + zchunk : g.mz;
  double vmax = -1.0e300;
  if (gidx < (long long)g.my * mx4) {
    const int y = (int)(gidx / mx4);
    const int x0 = (int)(gidx - (long long)y * mx4) << 2;
    const int gs = g.gs, w = gs + 1;
    const bool ys = (y == gs) || (y == g.my - gs - 1);
    const bool yi = y > gs && y < g.my - gs - 1;
    const bool yfull = y >= 1 && y < g.my - 1;
    const bool ring_y = (g.phys[2] && y < w) || (g.phys[3] && y >= g.my - w);
    const bool int_y = y >= gs && y < g.my - gs;
    // per-lane (x) masks, bit l = cell x0 + l
    unsigned mA = 0, mB = 0, mC = 0, ringx = 0, intx = 0;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const int x = x0 + l;
      const bool xs = (x == gs) || (x == g.mx - gs - 1);
      const bool xi = x > gs && x < g.mx - gs - 1;
      if ((xs && yfull) || (xi && ys)) mA |= 1u << l;
      if (xi && yi) mB |= 1u << l;
      if (yi && x >= 1 && x < g.mx - 1) mC |= 1u << l;
      if (ring_y || (g.phys[4] && x < w) || (g.phys[5] && x >= g.mx - w)) ringx |= 1u << l;
      if (int_y && x >= gs && x < g.mx - gs) intx |= 1u << l;
    }
    const long long n = g.vol, sy = g.mx, sz = g.plane;
    const T fsv[3] = {u0, u1, u2};
    long long i = (long long)zb * g.plane + (long long)y * g.mx + x0;
    for (int z = zb; z < ze; ++z, i += sz) {
      const bool zfull = z >= 1 && z < g.mz - 1;
      const bool zin = z >= gs && z < g.mz - gs;
      const bool zi = z > gs && z < g.mz - gs - 1;
      const bool ringz = (g.phys[0] && z < w) || (g.phys[1] && z >= g.mz - w);
      const unsigned ring = ringz ? 0xFu : ringx;
      const unsigned wr = (zfull ? mA : 0u) | (zin ? mB : 0u) | (zi ? mC : 0u);
      const unsigned calc = wr & ~ring;          // cells that get the curl
      const unsigned keep = ~(wr | ring) & 0xFu;  // cells that keep their old value (+ free stream)
      T c[3][4];
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int l = 0; l < 4; ++l) c[k][l] = T(0);
      if (calc) {
        // a computed cell has 1 <= y < my-1 and 1 <= z < mz-1: the y / z neighbour rows exist
        const T* fx = psi;
        const T* fy = psi + n;
        const T* fz = psi + 2 * n;
        const V fz_yp = *reinterpret_cast<const V*>(fz + i + sy), fz_ym = *reinterpret_cast<const V*>(fz + i - sy);
        const V fy_zp = *reinterpret_cast<const V*>(fy + i + sz), fy_zm = *reinterpret_cast<const V*>(fy + i - sz);
        const V fx_zp = *reinterpret_cast<const V*>(fx + i + sz), fx_zm = *reinterpret_cast<const V*>(fx + i - sz);
        const V fx_yp = *reinterpret_cast<const V*>(fx + i + sy), fx_ym = *reinterpret_cast<const V*>(fx + i - sy);
        const V fz_c = *reinterpret_cast<const V*>(fz + i), fy_c = *reinterpret_cast<const V*>(fy + i);
        // x neighbours: lanes 1..2 from the centre vectors, the outer two scalar (guarded at the row ends)
        T fz_row[6], fy_row[6];
        fz_row[0] = x0 > 0 ? fz[i - 1] : T(0);
        fy_row[0] = x0 > 0 ? fy[i - 1] : T(0);
        fz_row[5] = x0 + 4 < g.mx ? fz[i + 4] : T(0);
        fy_row[5] = x0 + 4 < g.mx ? fy[i + 4] : T(0);
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          fz_row[l + 1] = fz_c.v[l];
          fy_row[l + 1] = fy_c.v[l];
        }
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          if (calc & (1u << l)) {
            c[0][l] = p * (fz_yp.v[l] - fz_ym.v[l] - fy_zp.v[l] + fy_zm.v[l]);
            c[1][l] = p * (fx_zp.v[l] - fx_zm.v[l] - fz_row[l + 2] + fz_row[l]);
            c[2][l] = p * (fy_row[l + 2] - fy_row[l] - fx_yp.v[l] + fx_ym.v[l]);
          }
        }
      }
      if (keep) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const V old = *reinterpret_cast<const V*>(u + i + k * n);
#pragma unroll
          for (int l = 0; l < 4; ++l)
            if (keep & (1u << l)) c[k][l] = old.v[l];
        }
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        V o;
#pragma unroll
        for (int l = 0; l < 4; ++l) o.v[l] = c[k][l] = c[k][l] + fsv[k];
        *reinterpret_cast<V*>(u + i + k * n) = o;
        if (forcing) {
          V zero;
#pragma unroll
          for (int l = 0; l < 4; ++l) zero.v[l] = T(0);
          *reinterpret_cast<V*>(forcing + i + k * n) = zero;
        }
      }
      if (zin) {
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          if (intx & (1u << l)) {
            const T s = fabs(c[0][l]) + fabs(c[1][l]) + fabs(c[2][l]);
            vmax = (double)s > vmax ? (double)s : vmax;
          }
        }
      }
    }
  }
  if (max_out) {
    vmax = sb_block_reduce<true>(vmax);
    if (threadIdx.x == 0) atomicMax((unsigned long long*)max_out, sb_key_from_double(vmax));
  }
}

extern "C" int sb200_velocity_from_stream_function(const sb200_grid_t* gr, void* velocity,
                                                   const void* stream_func, double prefactor,
                                                   const double* free_stream, void* forcing, void* max_out,
                                                   void* stream) {
  SbGeom g;
  SB_REQUIRE(sb_make_geom(gr, &g) == 0, "bad grid");
  SB_REQUIRE(g.gs >= 1, "ghost_size < kernel_support");
  double fs[3] = {0, 0, 0};
  if (free_stream)
    for (int k = 0; k < g.dim; ++k) fs[k] = free_stream[k];
  if (max_out) {
    int e = sb_memset_async(max_out, 0, 8, stream);
    SB_REQUIRE(e == 0, "velocity: memset failed");
  }
  const int zchunk = g.mz >= 64 ? 16 : (g.mz >= 16 ? 8 : g.mz);
  dim3 block(256);
  dim3 grid((unsigned)((g.plane + 255) / 256), (unsigned)((g.mz + zchunk - 1) / zchunk));
  const size_t esz = gr->dtype == SB200_F32 ? 4 : 8;
  const bool aligned = (((uintptr_t)velocity | (uintptr_t)stream_func | (uintptr_t)forcing) % (4 * esz)) == 0;
  if (g.dim == 3 && g.mx % 4 == 0 && aligned) {
    // four x cells per thread, 16-byte (float) / 32-byte (double) accesses
    dim3 grid4((unsigned)(((long long)g.my * (g.mx / 4) + 255) / 256), grid.y);
    if (gr->dtype == SB200_F32) {
      SB_LAUNCH_COOP(sb_velocity_vec4_kernel<float>, grid4, block, 0, stream, g, (float*)velocity,
                     (const float*)stream_func, (float)prefactor, (float)fs[0], (float)fs[1], (float)fs[2],
                     (float*)forcing, max_out, zchunk);
    } else {
      SB_LAUNCH_COOP(sb_velocity_vec4_kernel<double>, grid4, block, 0, stream, g, (double*)velocity,
                     (const double*)stream_func, prefactor, fs[0], fs[1], fs[2], (double*)forcing, max_out,
                     zchunk);
    }
  } else if (gr->dtype == SB200_F32) {
    SB_LAUNCH_COOP(sb_velocity_kernel<float>, grid, block, 0, stream, g, (float*)velocity,
                   (const float*)stream_func, (float)prefactor, (float)fs[0], (float)fs[1], (float)fs[2],
                   (float*)forcing, max_out, zchunk);
  } else {
    SB_LAUNCH_COOP(sb_velocity_kernel<double>, grid, block, 0, stream, g, (double*)velocity,
                   (const double*)stream_func, prefactor, fs[0], fs[1], fs[2], (double*)forcing, max_out,
                   zchunk);
  }
  SB_CHECK_LAUNCH("velocity_from_stream_function");
  if (max_out) {
    SB_LAUNCH(sb_decode_key_kernel, dim3(1), dim3(32), 0, stream, max_out);
    SB_CHECK_LAUNCH("velocity_decode");
  }
  return 0;
}
