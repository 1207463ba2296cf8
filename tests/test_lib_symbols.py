"""The C-ABI library builds for sm_100a, loads, and exports every symbol that
include/sopht_b200.h declares (no compute calls: no GPU here)."""
import os
import re
import subprocess

import pytest

from sopht_mpi_b200 import _lib, build

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "sopht_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sb200_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    return build.build()


def test_header_declares_what_ctypes_binds():
    assert _header_symbols() == sorted(_lib.PROTOTYPES)


def test_library_exports_every_declared_symbol(lib_path):
    lib = _lib.bind(lib_path)  # raises AttributeError on a missing symbol
    assert lib.sb200_version() >= 100
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\sT\s+(sb200_[a-z0-9_]+)", out))
    assert set(_header_symbols()) <= exported


def test_library_contains_sm100a_code(lib_path):
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_product_loader_fails_loudly_without_library(tmp_path, monkeypatch):
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "missing.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(_lib.SophtB200Error):
        _lib.load()


def test_operators_refuse_to_run_without_cuda():
    import numpy as np
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sopht_mpi_b200.numeric.eulerian_grid_ops import gen_curl_pyst_mpi_kernel_3d
    from sopht_mpi_b200.utils import MPIConstruct3D, MPIGhostCommunicator3D

    mc = MPIConstruct3D(8, 8, 8, real_t=np.float32)
    gc = MPIGhostCommunicator3D(ghost_size=2, mpi_construct=mc)
    with pytest.raises(_lib.SophtB200Error):
        gen_curl_pyst_mpi_kernel_3d(real_t=np.float32, mpi_construct=mc, ghost_exchange_communicator=gc)
