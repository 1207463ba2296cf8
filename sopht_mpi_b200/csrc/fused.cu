// Fused vorticity update: one streaming pass over (omega, u) instead of the reference's
// cross product + curl update + three diffusion-flux/add sweeps
// (reference simulator/flow/flow_simulators_mpi_3d.py:395-411):
//     buf    = u x omega                      (every cell)
//     omega2 = omega + p * curl(buf)          (cells the MPI wrapper writes)
//     out    = omega2 + d * lap7(omega2)      (same cells, zero flux on the physical ring)
// 2.5D scheme: a CTA owns a (TY x TX) tile of the (y,x) plane and marches along z; the
// cross product is staged on the tile + 2 halo cells, omega2 on tile + 1, both in rings of
// three planes in shared memory, so every input plane is read from HBM once (plus the
// in-plane halo overlap) and the output written once: ~9 W per cell instead of 33 W.
#include "sb200_common.h"

template <int TY, int TX>
struct FusedTile {
  static constexpr int P2 = TX + 4, R2 = (TY + 4) * P2;  // tile + halo 2
  static constexpr int P1 = TX + 2, R1 = (TY + 2) * P1;  // tile + halo 1
  // floats: buf ring 3 planes x 3 comps, omega ring 2 x 3, omega2 ring 3 x 3
  static constexpr int ELEMS = 9 * R2 + 6 * R1 + 9 * R1;
};

// written-cell mask of a support-1 wrapper with the z conditions opened on faces that are
// NOT physical (inner slab faces): there the ghost planes hold exchanged data and the update
// applied to them reproduces what the neighbouring slab computes for its own interior.
SB_D bool sb_written_open(const SbGeom& g, int z, int y, int x) {
  const int gs = g.gs;
  const bool xs = (x == gs) || (x == g.mx - gs - 1);
  const bool xi = x > gs && x < g.mx - gs - 1;
  const bool ys = (y == gs) || (y == g.my - gs - 1);
  const bool yi = y > gs && y < g.my - gs - 1;
  const bool yfull = y >= 1 && y < g.my - 1;
  const bool zfull = z >= 1 && z < g.mz - 1;
  const int zlo = g.phys[0] ? gs : 1, zhi = g.phys[1] ? g.mz - gs : g.mz - 1;
  const bool zin = z >= zlo && z < zhi;
  return (xs && yfull && zfull) || (xi && ys && zfull) || (xi && yi && zin);
}

template <typename T, int TY, int TX, int NT>
__global__ void __launch_bounds__(NT)
    sb_vorticity_fused_kernel(SbGeom g, T* __restrict__ out, const T* __restrict__ w, const T* __restrict__ u,
                              T p, T d, int zchunk) {
  using FT = FusedTile<TY, TX>;
  SB_DYN_SMEM(smem_raw);
  T* sbuf = reinterpret_cast<T*>(smem_raw);  // [3][3][R2]
  T* sw1 = sbuf + 9 * FT::R2;                // [2][3][R1]
  T* sw2 = sw1 + 6 * FT::R1;                 // [3][3][R1]
  const int tid = threadIdx.x;
  const int y0 = blockIdx.y * TY, x0 = blockIdx.x * TX;
  const int zb = blockIdx.z * zchunk;
  const int ze = zb + zchunk < g.mz ? zb + zchunk : g.mz;
  const long long vol = g.vol;

  for (int zf = zb - 2; zf <= ze + 1; ++zf) {
    // ---- A: cross product of plane zf on tile + 2, omega of plane zf on tile + 1
    if (zf >= 0 && zf < g.mz) {
      T* b = sbuf + (zf % 3) * 3 * FT::R2;
      T* s1 = sw1 + (zf & 1) * 3 * FT::R1;
      for (int i = tid; i < FT::R2; i += NT) {
        const int ry = i / FT::P2, rx = i - ry * FT::P2;
        const int y = y0 + ry - 2, x = x0 + rx - 2;
        T w0 = 0, w1 = 0, w2 = 0, u0 = 0, u1 = 0, u2 = 0;
        if (y >= 0 && y < g.my && x >= 0 && x < g.mx) {
          const long long gi = g.idx(zf, y, x);
          w0 = w[gi];
          w1 = w[gi + vol];
          w2 = w[gi + 2 * vol];
          u0 = u[gi];
          u1 = u[gi + vol];
          u2 = u[gi + 2 * vol];
        }
        b[i] = u1 * w2 - u2 * w1;
        b[FT::R2 + i] = u2 * w0 - u0 * w2;
        b[2 * FT::R2 + i] = u0 * w1 - u1 * w0;
        if (ry >= 1 && ry <= TY + 2 && rx >= 1 && rx <= TX + 2) {
          const int j = (ry - 1) * FT::P1 + (rx - 1);
          s1[j] = w0;
          s1[FT::R1 + j] = w1;
          s1[2 * FT::R1 + j] = w2;
        }
      }
    }
    __syncthreads();
    // ---- B: omega2 of plane zc = zf - 1 on tile + 1
    const int zc = zf - 1;
    if (zc >= 0 && zc < g.mz) {
      const T* bm = sbuf + ((zc + 2) % 3) * 3 * FT::R2;  // plane zc - 1
      const T* b0 = sbuf + (zc % 3) * 3 * FT::R2;
      const T* bp = sbuf + ((zc + 1) % 3) * 3 * FT::R2;
      const T* s1 = sw1 + (zc & 1) * 3 * FT::R1;
      T* s2 = sw2 + (zc % 3) * 3 * FT::R1;
      for (int j = tid; j < FT::R1; j += NT) {
        const int ry = j / FT::P1, rx = j - ry * FT::P1;
        const int y = y0 + ry - 1, x = x0 + rx - 1;
        T c0 = s1[j], c1 = s1[FT::R1 + j], c2 = s1[2 * FT::R1 + j];
        if (y >= 0 && y < g.my && x >= 0 && x < g.mx && sb_written_open(g, zc, y, x)) {
          const int i = (ry + 1) * FT::P2 + (rx + 1);  // same cell in the tile + 2 frame
          const T* bx0 = b0;
          const T* by0 = b0 + FT::R2;
          const T* bz0 = b0 + 2 * FT::R2;
          c0 += p * (bz0[i + FT::P2] - bz0[i - FT::P2] - bp[FT::R2 + i] + bm[FT::R2 + i]);
          c1 += p * (bp[i] - bm[i] - bz0[i + 1] + bz0[i - 1]);
          c2 += p * (by0[i + 1] - by0[i - 1] - bx0[i + FT::P2] + bx0[i - FT::P2]);
        }
        s2[j] = c0;
        s2[FT::R1 + j] = c1;
        s2[2 * FT::R1 + j] = c2;
      }
    }
    __syncthreads();
    // ---- C: diffusion of plane zo = zf - 2 on the tile, write out
    const int zo = zf - 2;
    if (zo >= zb && zo < ze) {
      const T* qm = sw2 + ((zo + 2) % 3) * 3 * FT::R1;
      const T* q0 = sw2 + (zo % 3) * 3 * FT::R1;
      const T* qp = sw2 + ((zo + 1) % 3) * 3 * FT::R1;
      for (int k = tid; k < TY * TX; k += NT) {
        const int ty = k / TX, tx = k - ty * TX;
        const int y = y0 + ty, x = x0 + tx;
        if (y < g.my && x < g.mx) {
          const int j = (ty + 1) * FT::P1 + (tx + 1);
          const bool lap = !g.in_ring(zo, y, x) && sb_written_open(g, zo, y, x);
          const long long gi = g.idx(zo, y, x);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const T* q = q0 + c * FT::R1;
            T v = q[j];
            if (lap) {
              const T s = q[j + 1] + q[j - 1] + q[j + FT::P1] + q[j - FT::P1] + qp[c * FT::R1 + j] +
                          qm[c * FT::R1 + j];
              v += d * (s - T(6) * q[j]);
            }
            out[gi + c * vol] = v;
          }
        }
      }
    }
    // (the barrier after phase A of the next iteration orders C against the next B)
  }
}

template <typename T, int TY, int TX, int NT>
static int launch_fused(const SbGeom& g, void* out, const void* w, const void* u, double p, double d,
                        void* stream) {
  using FT = FusedTile<TY, TX>;
  const size_t smem = sizeof(T) * FT::ELEMS;
  const unsigned gx = (g.mx + TX - 1) / TX, gy = (g.my + TY - 1) / TY;
  // enough z chunks for ~2 waves of 148 SMs, each chunk at least 16 planes
  int chunks = (int)((2 * 148 + gx * gy - 1) / (gx * gy));
  if (chunks < 1) chunks = 1;
  int zchunk = (g.mz + chunks - 1) / chunks;
  if (zchunk < 16) zchunk = 16;
  if (zchunk > g.mz) zchunk = g.mz;
  chunks = (g.mz + zchunk - 1) / zchunk;
  SB_KERNEL_ATTR_SMEM((sb_vorticity_fused_kernel<T, TY, TX, NT>), smem);
  SB_LAUNCH_COOP((sb_vorticity_fused_kernel<T, TY, TX, NT>), dim3(gx, gy, (unsigned)chunks), dim3(NT), smem,
                 stream, g, (T*)out, (const T*)w, (const T*)u, (T)p, (T)d, zchunk);
  SB_CHECK_LAUNCH("vorticity_fused");
  return 0;
}

extern "C" int sb200_vorticity_rhs_fused_3d(const sb200_grid_t* gr, void* out, const void* vorticity,
                                            const void* velocity, const void* forcing,
                                            double curl_prefactor, double nu_dt_by_dx2, void* stream) {
  SbGeom g;
  SB_REQUIRE(sb_make_geom(gr, &g) == 0, "bad grid");
  SB_REQUIRE(g.dim == 3, "vorticity_rhs_fused_3d: 3D only");
  SB_REQUIRE(g.gs >= 2, "vorticity_rhs_fused_3d needs ghost_size >= 2");
  SB_REQUIRE(out && vorticity && velocity && out != vorticity, "vorticity_rhs_fused_3d: bad pointers");
  SB_REQUIRE(forcing == nullptr,
             "vorticity_rhs_fused_3d: apply the forcing update first "
             "(sb200_update_vorticity_from_velocity_forcing)");
  if (gr->dtype == SB200_F32)
    return launch_fused<float, 16, 64, 512>(g, out, vorticity, velocity, curl_prefactor, nu_dt_by_dx2, stream);
  return launch_fused<double, 8, 64, 512>(g, out, vorticity, velocity, curl_prefactor, nu_dt_by_dx2, stream);
}
