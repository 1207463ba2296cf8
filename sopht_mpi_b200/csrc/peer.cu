// NVLink peer copies for the transposes of the distributed Poisson solve: the destination is
// another rank's exchange buffer, mapped into this process through CUDA IPC by the Python side
// (utils/peer.py).  cudaMemcpyPeerAsync runs on the copy engines and goes straight over NVLink
// once peer access is enabled between the two devices.
#include "sb200_common.h"

extern "C" int sb200_enable_peer_access(int device, int peer_device) {
#ifndef SB200_EMU
  if (device == peer_device) return 0;
  int can = 0;
  SB_REQUIRE(cudaDeviceCanAccessPeer(&can, device, peer_device) == cudaSuccess, "peer access: query failed");
  SB_REQUIRE(can, "peer access: the two devices cannot access each other's memory");
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(device);
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  cudaSetDevice(cur);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();  // clear the sticky-less error state
    return 0;
  }
  SB_REQUIRE(e == cudaSuccess, "peer access: cudaDeviceEnablePeerAccess failed");
#else
  (void)device; (void)peer_device;
#endif
  return 0;
}

extern "C" int sb200_peer_copy(void* dst, int dst_device, const void* src, int src_device, int64_t bytes,
                               void* stream) {
  SB_REQUIRE(dst && src && bytes >= 0, "peer_copy: bad arguments");
#ifndef SB200_EMU
  const cudaError_t e = cudaMemcpyPeerAsync(dst, dst_device, src, src_device, (size_t)bytes, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    sb_set_error("peer_copy: %s", cudaGetErrorString(e));
    return -2;
  }
#else
  (void)dst_device; (void)src_device; (void)stream;
  memcpy(dst, src, (size_t)bytes);
#endif
  return 0;
}

// all blocks of one exchange in a single call: block k goes to dst[k] on device dst_device[k]
extern "C" int sb200_peer_copy_blocks(int n, void* const* dst, const int* dst_device, const void* const* src,
                                      int src_device, int64_t bytes, void* stream) {
  SB_REQUIRE(n >= 0 && dst && dst_device && src, "peer_copy_blocks: bad arguments");
  for (int k = 0; k < n; ++k) {
    const int e = sb200_peer_copy(dst[k], dst_device[k], src[k], src_device, bytes, stream);
    if (e) return e;
  }
  return 0;
}

// The same exchange as ONE kernel: every destination block is written through its IPC-mapped pointer
// by its own group of thread blocks (16-byte loads from the local send buffer, 16-byte stores that
// travel over NVLink), so the copies to the P - 1 peers run concurrently through the NVSwitch instead
// of one after the other on a copy engine, and no foreign device context is involved (the mapped
// pointers are peer-addressable under UVA).  The grid is kept small (blocks_per_peer x n blocks):
// the stores are NVLink-bound, and the remaining SMs stay available to the transform kernels of the
// neighbouring component that the exchange overlaps.
struct SbPushPlan {
  void* dst[8];
  const void* src[8];
  int n;
};
__global__ void __launch_bounds__(512) sb_push_blocks_kernel(SbPushPlan plan, long long count16) {
  const int k = blockIdx.y;
  const float4* __restrict__ s = reinterpret_cast<const float4*>(plan.src[k]);
  float4* __restrict__ d = reinterpret_cast<float4*>(plan.dst[k]);
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // four independent 16-byte transfers in flight per thread
  for (; i + 3 * stride < count16; i += 4 * stride) {
    const float4 a = s[i], b = s[i + stride], c = s[i + 2 * stride], e = s[i + 3 * stride];
    d[i] = a;
    d[i + stride] = b;
    d[i + 2 * stride] = c;
    d[i + 3 * stride] = e;
  }
  for (; i < count16; i += stride) d[i] = s[i];
}

extern "C" int sb200_peer_push_blocks(int n, void* const* dst, const void* const* src, int64_t bytes,
                                      int blocks_per_peer, void* stream) {
  SB_REQUIRE(n >= 0 && n <= 8 && dst && src, "peer_push_blocks: bad arguments (at most 8 blocks)");
  SB_REQUIRE(bytes % 16 == 0, "peer_push_blocks: block size must be a multiple of 16 bytes");
  if (n == 0 || bytes == 0) return 0;
  SbPushPlan plan;
  plan.n = n;
  for (int k = 0; k < n; ++k) {
    plan.dst[k] = dst[k];
    plan.src[k] = src[k];
  }
  if (blocks_per_peer <= 0) blocks_per_peer = 8;
  SB_LAUNCH(sb_push_blocks_kernel, dim3((unsigned)blocks_per_peer, (unsigned)n), dim3(512), 0, stream, plan,
            (long long)(bytes / 16));
  SB_CHECK_LAUNCH("peer_push_blocks");
  return 0;
}
