// Micro-benchmark: issue throughput of scalar FFMA/FADD against the packed sm_100 forms
// (fma.rn.f32x2 / add.rn.f32x2 -> FFMA2 / FADD2).  Prints Gop/s (thread-level lane operations).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

template <int MODE>
__global__ void k(float* out, int iters, float s) {
  float a[16];
  u64 p[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = ((u64)__float_as_uint(a[2 * i]) << 32) | __float_as_uint(a[2 * i + 1]);
  const u64 ps = ((u64)__float_as_uint(s) << 32) | __float_as_uint(s);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // 16 scalar FFMA
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], s, a[(i + 1) & 15]);
    } else if (MODE == 1) {  // 8 FFMA2 (same lane work as mode 0)
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], ps, p[(i + 1) & 7]);
    } else if (MODE == 2) {  // 16 scalar FADD
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = a[i] + a[(i + 1) & 15];
    } else if (MODE == 3) {  // 8 FADD2
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = add2(p[i], p[(i + 1) & 7]);
    } else if (MODE == 4) {  // 8 FMUL2
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = mul2(p[i], ps);
    } else if (MODE == 5) {  // mixed: 8 FADD2 + 8 scalar FFMA (do they dual-issue / share the pipe?)
#pragma unroll
      for (int i = 0; i < 8; ++i) { p[i] = add2(p[i], p[(i + 1) & 7]); a[i] = fmaf(a[i], s, a[(i + 1) & 7]); }
    }
  }
  float r = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += a[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) r += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, double lanes_per_iter) {
  float* out;
  const int blocks = 148 * 8, threads = 256, iters = 4096;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, 16, 1.0001f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, iters, 1.0001f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)blocks * threads * iters * lanes_per_iter;
  printf("%-34s %8.3f ms  %9.1f G lane-ops/s   %8.1f G warp-inst/s\n", name, ms, ops / ms * 1e-6,
         ops / ms * 1e-6 / 32 / (MODE == 1 || MODE == 3 || MODE == 4 ? 2 : MODE == 5 ? 24.0 / 16 : 1));
  cudaFree(out);
}
int main() {
  run<0>("FFMA scalar x16", 16);
  run<1>("FFMA2 x8 (16 lanes)", 16);
  run<2>("FADD scalar x16", 16);
  run<3>("FADD2 x8 (16 lanes)", 16);
  run<4>("FMUL2 x8 (16 lanes)", 16);
  run<5>("8 FADD2 + 8 FFMA (24 lanes)", 24);
  return 0;
}
