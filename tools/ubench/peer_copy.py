"""Bandwidth of CUDA-IPC peer copies between two ranks (torchrun --nproc-per-node 2)."""
import ctypes
import os
import sys

import torch
import torch.distributed as dist
from torch.multiprocessing.reductions import reduce_tensor

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from sopht_mpi_b200 import _lib  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl")
    g = dist.new_group(backend="gloo")
    lib = _lib.load()
    n = 64 << 20
    mine = torch.zeros(n // 4, dtype=torch.float32, device="cuda")
    src = torch.ones(n // 4, dtype=torch.float32, device="cuda")
    handles = [None, None]
    dist.all_gather_object(handles, reduce_tensor(mine), group=g)
    fn, args = handles[1 - rank]
    peer = fn(*args)
    print(rank, "peer tensor device", peer.device, "can access", torch.cuda.can_device_access_peer(rank, 1 - rank), flush=True)
    for a, b in ((rank, 1 - rank), (1 - rank, rank)):
        print(rank, "enable", a, b, lib.sb200_enable_peer_access(a, b), flush=True)
    # the library's own allocation + IPC mapping (opened with THIS rank's device current)
    own = ctypes.c_void_p()
    _lib.check(lib, lib.sb200_peer_alloc(n, ctypes.byref(own)))
    hbuf = ctypes.create_string_buffer(64)
    _lib.check(lib, lib.sb200_ipc_export(own, hbuf))
    raws = [None, None]
    dist.all_gather_object(raws, hbuf.raw, group=g)
    peer_own = ctypes.c_void_p()
    _lib.check(lib, lib.sb200_ipc_open(ctypes.create_string_buffer(raws[1 - rank], 64), ctypes.byref(peer_own)))
    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)

    def timed(label, f):
        f()
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            f()
        b.record()
        torch.cuda.synchronize()
        print(rank, label, f"{n * 10 / (a.elapsed_time(b) * 1e-3) / 1e9:.1f} GB/s", flush=True)
        dist.barrier()

    timed("push cudaMemcpyPeerAsync", lambda: lib.sb200_peer_copy(
        ctypes.c_void_p(peer.data_ptr()), 1 - rank, ctypes.c_void_p(src.data_ptr()), rank, n, sp))
    timed("pull cudaMemcpyPeerAsync", lambda: lib.sb200_peer_copy(
        ctypes.c_void_p(src.data_ptr()), rank, ctypes.c_void_p(peer.data_ptr()), 1 - rank, n, sp))
    timed("torch copy_ push", lambda: peer.copy_(src, non_blocking=True))
    for bpp in (4, 8, 16, 32, 64):
        dp = (ctypes.c_void_p * 1)(peer_own.value)
        spp = (ctypes.c_void_p * 1)(src.data_ptr())
        timed(f"push kernel, own IPC mapping ({bpp} blocks of 512 threads)", lambda: lib.sb200_peer_push_blocks(
            1, dp, spp, n, bpp, None, 0, None, sp))
    check = torch.as_tensor(type("A", (), {"__cuda_array_interface__": {
        "shape": (n // 4,), "typestr": "<f4", "data": (own.value, False), "version": 2}})(), device="cuda")
    torch.cuda.synchronize()
    dist.barrier()
    print(rank, "own buffer filled by the peer:", bool((check == 1).all().item()), flush=True)
    other = torch.empty_like(src)
    timed("nccl send/recv", lambda: dist.batch_isend_irecv(
        [dist.P2POp(dist.isend, src, 1 - rank), dist.P2POp(dist.irecv, other, 1 - rank)])[-1].wait())
    dist.barrier()
    del peer
    lib.sb200_ipc_close(peer_own)
    dist.barrier()
    lib.sb200_peer_free(own)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
