"""Local stages of the z-slab Poisson solve of ONE rank, timed on a single GPU (no exchange):
    python tools/bench_slab_local.py NZ NY NX NRANKS"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from sopht_mpi_b200 import _lib  # noqa: E402


def main():
    nz, ny, nx, nranks = (int(a) for a in sys.argv[1:5])
    lib = _lib.load()
    gs = 2
    h = ctypes.c_void_p()
    _lib.check(lib, lib.sb200_poisson_create(ctypes.byref(h), 3, _lib.F32, nz, ny, nx, gs, 1.0, 0, nranks, 1, None))
    nfloat = int(lib.sb200_poisson_slab_buffer_bytes(h, 3)) // 4
    send = torch.zeros(nfloat, device="cuda")
    recv = torch.rand(nfloat, device="cuda")
    nzl = nz // nranks
    rhs = torch.rand((3, nzl + 2 * gs, ny + 2 * gs, nx + 2 * gs), device="cuda")
    sol = torch.zeros_like(rhs)
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    stages = {
        "slab_forward (x r2c, blocked rows)": lambda: lib.sb200_poisson_slab_forward(h, p(rhs), 3, p(send), None),
        "slab_spectral (y fwd, z fused, y inv)": lambda: lib.sb200_poisson_slab_spectral(h, p(recv), 3, None),
        "slab_backward (x c2r, blocked rows)": lambda: lib.sb200_poisson_slab_backward(h, p(sol), 3, p(send), None),
    }
    for name, fn in stages.items():
        for _ in range(2):
            _lib.check(lib, fn())
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        b.record()
        torch.cuda.synchronize()
        print(f"{name}: {a.elapsed_time(b) / 10:.3f} ms", flush=True)
    lib.sb200_poisson_destroy(h)


if __name__ == "__main__":
    main()
