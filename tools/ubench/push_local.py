"""The push kernel of the peer exchange on ONE device (local source and destination blocks): checks the
copy, the epoch flags and the wait kernel, and reports the device-local copy rate."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from sopht_mpi_b200 import _lib  # noqa: E402


def main():
    lib = _lib.load()
    nblk, n = 3, 32 << 20
    src = [torch.rand(n // 4, device="cuda") for _ in range(nblk)]
    dst = [torch.zeros(n // 4, device="cuda") for _ in range(nblk)]
    flags = torch.zeros(nblk, dtype=torch.int32, device="cuda")
    done = torch.zeros(8, dtype=torch.int32, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    dp = (ctypes.c_void_p * nblk)(*[t.data_ptr() for t in dst])
    sp = (ctypes.c_void_p * nblk)(*[t.data_ptr() for t in src])
    fp = (ctypes.c_void_p * nblk)(*[flags[k:k + 1].data_ptr() for k in range(nblk)])
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib, lib.sb200_peer_push_blocks(nblk, dp, sp, n, 8, None, 0, None, stream))
    torch.cuda.synchronize()
    print("plain push ok:", all(torch.equal(a, b) for a, b in zip(src, dst)), flush=True)
    for t in dst:
        t.zero_()
    for epoch in (1, 2):
        _lib.check(lib, lib.sb200_peer_push_blocks(nblk, dp, sp, n, 8, fp, epoch, ctypes.c_void_p(done.data_ptr()),
                                                   stream))
        _lib.check(lib, lib.sb200_peer_wait_flags(ctypes.c_void_p(flags.data_ptr()), nblk, epoch,
                                                  ctypes.c_void_p(err.data_ptr()), stream))
        torch.cuda.synchronize()
        print("epoch", epoch, "flags", flags.tolist(), "done", done.tolist()[:nblk], "err", int(err.item()),
              "data ok:", all(torch.equal(a, b) for a, b in zip(src, dst)), flush=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        lib.sb200_peer_push_blocks(nblk, dp, sp, n, 16, None, 0, None, stream)
    b.record()
    torch.cuda.synchronize()
    print(f"local copy rate {nblk * n * 10 / (a.elapsed_time(b) * 1e-3) / 1e9:.0f} GB/s (read + write = 2x)", flush=True)


if __name__ == "__main__":
    main()
