"""Build the host-thread emulation of the CUDA kernels (TEST INFRASTRUCTURE ONLY).

Compiles sopht_mpi_b200/csrc/*.cu with g++ -DSB200_EMU into tests/emu/libsb200_emu.so so
that kernel logic can be checked against the oracle in the GPU-less container.
The product package never loads this library.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
CSRC = os.path.join(ROOT, "sopht_mpi_b200", "csrc")
LIB = os.path.join(HERE, "libsb200_emu.so")
SOURCES = ["stencils.cu", "reduce.cu", "ib.cu", "poisson.cu", "poisson_cufft.cu", "poisson_fft.cu",
           "fused.cu", "peer.cu"]


def build(force=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(HERE, "emu_rt.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")]
    deps.append(os.path.join(ROOT, "include", "sopht_b200.h"))
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(map(os.path.getmtime, deps)):
        return LIB
    cmd = ["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-DSB200_EMU=1",
           "-ffp-contract=off", "-I", CSRC, "-o", LIB]
    for s in srcs:
        cmd += ["-x", "c++", s]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emu build failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
