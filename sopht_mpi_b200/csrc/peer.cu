// NVLink peer copies for the transposes of the distributed Poisson solve: the destination is
// another rank's exchange buffer, mapped into this process through CUDA IPC by the Python side
// (utils/peer.py).  cudaMemcpyPeerAsync runs on the copy engines and goes straight over NVLink
// once peer access is enabled between the two devices.
#include "sb200_common.h"

#include <cstdlib>
#include <cstring>

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
  int* flag[8];  // where to signal "block k has landed" on its destination (NULL: no signalling)
  int n;
};
#ifdef SB200_EMU
static inline void __threadfence_system() {}
#endif
__global__ void __launch_bounds__(512)
    sb_push_blocks_kernel(SbPushPlan plan, long long count16, int epoch, int* done) {
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
  if (plan.flag[k] == nullptr) return;
  // signal the destination once ALL thread blocks of this destination have stored their share:
  // every thread makes its stores visible system-wide, the last block to arrive raises the flag
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int arrived = atomicAdd(&done[k], 1);
    if (arrived == (int)gridDim.x - 1) {
      done[k] = 0;
      __threadfence_system();
      *reinterpret_cast<volatile int*>(plan.flag[k]) = epoch;
    }
  }
}

// wait (on the stream) until the n flags of this rank's buffer have reached `epoch`: every source rank
// has finished writing its block.  A bounded spin: after ~2 s without progress the kernel gives up and
// raises *error instead of hanging the device.
__global__ void sb_wait_flags_kernel(const int* flags, int n, int epoch, int* error, long long limit_ticks) {
  const int k = threadIdx.x;
  if (k >= n) return;
#ifndef SB200_EMU
  const volatile int* f = flags + k;
  const long long t0 = clock64();
  while (*f < epoch) {
    if (clock64() - t0 > limit_ticks) {
      *error = 1;
      break;
    }
    __nanosleep(64);
  }
  __threadfence_system();
#else
  if (flags[k] < epoch) *error = 1;
#endif
}

extern "C" int sb200_peer_push_blocks(int n, void* const* dst, const void* const* src, int64_t bytes,
                                      int blocks_per_peer, void* const* signal_flags, int epoch, void* done_counters,
                                      void* stream) {
  SB_REQUIRE(n >= 0 && n <= 8 && dst && src, "peer_push_blocks: bad arguments (at most 8 blocks)");
  SB_REQUIRE(bytes % 16 == 0, "peer_push_blocks: block size must be a multiple of 16 bytes");
  SB_REQUIRE(signal_flags == nullptr || done_counters != nullptr, "peer_push_blocks: signalling needs the counters");
  if (n == 0 || bytes == 0) return 0;
  SbPushPlan plan;
  plan.n = n;
  for (int k = 0; k < 8; ++k) {
    plan.dst[k] = k < n ? dst[k] : nullptr;
    plan.src[k] = k < n ? src[k] : nullptr;
    plan.flag[k] = (k < n && signal_flags) ? (int*)signal_flags[k] : nullptr;
  }
  if (blocks_per_peer <= 0) blocks_per_peer = 8;
  SB_LAUNCH_COOP(sb_push_blocks_kernel, dim3((unsigned)blocks_per_peer, (unsigned)n), dim3(512), 0, stream, plan,
                 (long long)(bytes / 16), epoch, (int*)done_counters);
  SB_CHECK_LAUNCH("peer_push_blocks");
  return 0;
}

extern "C" int sb200_peer_wait_flags(const void* flags, int n, int epoch, void* error_flag, void* stream) {
  SB_REQUIRE(flags && error_flag && n >= 1 && n <= 32, "peer_wait_flags: bad arguments");
  // ~2 s by default; SB200_PEER_WAIT_S raises it (ranks that time-share ONE device, as in
  // tests/test_gpu_ranks_on_one_device.py, hand the GPU to each other one time slice at a time)
  static const long long limit = (long long)(2.0e9 * (getenv("SB200_PEER_WAIT_S") ? atof(getenv("SB200_PEER_WAIT_S")) : 2.0));
  SB_LAUNCH(sb_wait_flags_kernel, dim3(1), dim3(32), 0, stream, (const int*)flags, n, epoch, (int*)error_flag, limit);
  SB_CHECK_LAUNCH("peer_wait_flags");
  return 0;
}

// ---- exchange buffers with our own CUDA IPC mapping.  The buffers are plain cudaMalloc allocations
// (so the pointer IS the base the handle refers to); a peer opens the handle while ITS OWN device is
// current, with cudaIpcMemLazyEnablePeerAccess: the mapping lives in the peer's compute context and its
// kernels can store through it over NVLink.  (torch's tensor sharing maps imported storages in a
// context on the EXPORTING device instead, which kernel stores from the importing device fault on.)
extern "C" int sb200_peer_alloc(int64_t bytes, void** ptr_out) {
  SB_REQUIRE(ptr_out && bytes > 0, "peer_alloc: bad arguments");
#ifndef SB200_EMU
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)bytes);
  // cudaMemset of device memory only ENQUEUES the fill: without this, a peer that maps the buffer can
  // raise an epoch flag in it before the fill runs and have it wiped (seen as a lost halo signal with
  // several ranks on one device, tests/test_gpu_ranks_on_one_device.py)
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    sb_set_error("peer_alloc: %s", cudaGetErrorString(e));
    return -2;
  }
  *ptr_out = p;
#else
  *ptr_out = calloc(1, (size_t)bytes);
#endif
  return 0;
}
extern "C" int sb200_peer_free(void* ptr) {
#ifndef SB200_EMU
  if (ptr) cudaFree(ptr);
#else
  free(ptr);
#endif
  return 0;
}
extern "C" int sb200_ipc_export(const void* ptr, void* handle_out_64) {
  SB_REQUIRE(ptr && handle_out_64, "ipc_export: bad arguments");
#ifndef SB200_EMU
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(ptr));
  if (e != cudaSuccess) {
    sb_set_error("ipc_export: %s", cudaGetErrorString(e));
    return -2;
  }
  memcpy(handle_out_64, &h, 64);
#else
  memset(handle_out_64, 0, 64);
#endif
  return 0;
}
extern "C" int sb200_ipc_open(const void* handle_64, void** ptr_out) {
  SB_REQUIRE(handle_64 && ptr_out, "ipc_open: bad arguments");
#ifndef SB200_EMU
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_64, 64);
  void* p = nullptr;
  const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    sb_set_error("ipc_open: %s", cudaGetErrorString(e));
    return -2;
  }
  *ptr_out = p;
#else
  *ptr_out = nullptr;
#endif
  return 0;
}
extern "C" int sb200_ipc_close(void* ptr) {
#ifndef SB200_EMU
  if (ptr) cudaIpcCloseMemHandle(ptr);
#else
  (void)ptr;
#endif
  return 0;
}
