// Host-thread emulation runtime for the kernels in sopht_mpi_b200/csrc
// (TEST INFRASTRUCTURE ONLY; see sb200_rt.h).  Blocks run one after another;
// "coop" launches give every CUDA thread an OS thread so that __syncthreads and
// warp shuffles work; plain launches run the threads of a block sequentially.
#define SB200_EMU 1
#include "sb200_rt.h"

namespace sbemu {
thread_local uint3 t_threadIdx, t_blockIdx;
thread_local dim3 t_blockDim, t_gridDim;
thread_local BlockCtx* t_ctx = nullptr;
thread_local unsigned t_linear_tid = 0;

void launch(dim3 grid, dim3 block, size_t smem, bool coop, const std::function<void()>& body) {
  const unsigned nthreads = block.x * block.y * block.z;
  BlockCtx ctx;
  ctx.dyn_smem.resize(smem + 16);
  ctx.xchg.resize(nthreads);
  if (coop) {
    ctx.block_bar = std::make_unique<std::barrier<>>(nthreads);
    for (unsigned w = 0; w < (nthreads + 31) / 32; ++w) {
      unsigned cnt = std::min(32u, nthreads - w * 32);
      ctx.warp_bar.push_back(std::make_unique<std::barrier<>>(cnt));
    }
  }
  auto run_thread = [&](unsigned bx, unsigned by, unsigned bz, unsigned tid) {
    t_ctx = &ctx;
    t_linear_tid = tid;
    t_blockDim = block;
    t_gridDim = grid;
    t_blockIdx = uint3{bx, by, bz};
    t_threadIdx = uint3{tid % block.x, (tid / block.x) % block.y, tid / (block.x * block.y)};
    body();
  };
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        if (!coop) {
          for (unsigned tid = 0; tid < nthreads; ++tid) run_thread(bx, by, bz, tid);
        } else {
          std::vector<std::thread> th;
          th.reserve(nthreads);
          for (unsigned tid = 0; tid < nthreads; ++tid)
            th.emplace_back(run_thread, bx, by, bz, tid);
          for (auto& t : th) t.join();
        }
      }
}
}  // namespace sbemu
