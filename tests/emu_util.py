"""Helpers for the emulation tests: call the C ABI on numpy (host) buffers through
the host-thread emulation build of the CUDA kernels (tests/emu)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from sopht_mpi_b200 import _lib  # noqa: E402

_emu = None


def emu():
    global _emu
    if _emu is None:
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import build_emu

        _emu = _lib.bind(build_emu.build())
    return _emu


def ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def call(fn_name, *args):
    lib = emu()
    err = getattr(lib, fn_name)(*args)
    _lib.check(lib, err)
