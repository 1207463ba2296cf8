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
