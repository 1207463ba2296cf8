"""All-to-all of equal blocks over NVLink peer memory (one process per GPU, one node).

Every rank allocates its exchange buffers (plain device allocations of ``libsophtb200``), publishes
their CUDA IPC handles and opens the handles of all other ranks WHILE ITS OWN DEVICE IS CURRENT, so the
mappings live in its compute context; an exchange is then ONE small kernel
(``sb200_peer_push_blocks``) whose thread blocks store every destination block straight into the
destination rank's buffer, all peers concurrently through the NVSwitch.  "All blocks have landed" is
signalled through peer memory too: the last thread block of every destination raises an epoch flag in
the destination's memory, and a one-warp kernel on the destination's stream waits for the P flags
(bounded spin), so no collective sits on the critical path.  Replaces the MPI ``Alltoallw`` inside
mpi4py-fft's transposes (reference ``poisson_solver_3d/fft_mpi_3d.py:27-48``).

``SB200_EXCHANGE`` selects the transport for measurements: ``push`` (default), ``push-nccl`` (push kernel +
a one-element NCCL all-reduce as the barrier), ``copy`` (one ``cudaMemcpyPeerAsync`` per destination on
the stream + all-reduce, round 1), ``nccl`` (all-to-all).  If CUDA IPC is not available (different
nodes, no peer access) the exchange falls back to NCCL.
"""
import ctypes
import os

import torch
import torch.distributed as dist

from .. import _lib
from .comm import host_group
from .device import current_stream_ptr
from .logger import logger


class _RawDeviceArray:
    """``__cuda_array_interface__`` view of a raw device allocation (so torch can alias it)"""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


class PeerExchange:
    def __init__(self, n_buffers, n_float32, device, rank, nranks, use_peer_copies=True):
        self.rank, self.nranks, self.device = rank, nranks, device
        self.n_buffers, self.n_float32 = n_buffers, int(n_float32)
        self.transport = os.environ.get("SB200_EXCHANGE", "push")
        # thread blocks per destination: ~32 blocks of 512 threads saturate one NVLink port pair
        # (tools/ubench/peer_copy.py: 8 -> 364, 16 -> 630, 32 -> 658 GB/s); with P - 1 destinations in
        # flight the egress limit is shared, so fewer blocks per destination leave more SMs to the
        # transform kernels the exchange overlaps
        default_blocks = max(4, -(-40 // max(nranks - 1, 1)))
        self.blocks_per_peer = int(os.environ.get("SB200_PUSH_BLOCKS", "0")) or default_blocks
        if self.transport == "nccl":
            use_peer_copies = False
        self._lib = _lib.load() if device.type == "cuda" else None
        self._raw = []          # device allocations owned by this object (sb200_peer_alloc)
        self._opened = []       # IPC mappings of the other ranks' allocations
        self.peer_ptr = None    # [rank][buffer] -> base address of that rank's buffer
        self.peer_flag_ptr = None
        self._plans = {}
        self._epoch = [0] * n_buffers
        want_ipc = use_peer_copies and nranks > 1 and device.type == "cuda"
        if want_ipc:
            # zero-initialised (the padding bins of the last kx block are never written)
            self.local = [self._alloc(n_float32 * 4, n_float32, "<f4") for _ in range(n_buffers)]
            # epoch flags [buffer][source rank] (written by the peers)
            self.flags = self._alloc(4 * n_buffers * nranks, n_buffers * nranks, "<i4").view(n_buffers, nranks)
        else:
            self.local = [torch.zeros(n_float32, dtype=torch.float32, device=device) for _ in range(n_buffers)]
            self.flags = torch.zeros((n_buffers, max(nranks, 1)), dtype=torch.int32, device=device)
        self._flag = torch.zeros(1, dtype=torch.float32, device=device)
        self._done = torch.zeros((n_buffers, 8), dtype=torch.int32, device=device)  # arrival counters
        self._err = torch.zeros(1, dtype=torch.int32, device=device)
        ok = 1.0
        if want_ipc:
            try:
                self._map_peers()
            except Exception as exc:  # pragma: no cover - depends on the node
                logger.warning(f"CUDA IPC peer mapping unavailable ({type(exc).__name__}: {exc}); "
                               "the transposes use NCCL all-to-all")
                self.peer_ptr = None
                ok = 0.0
        flag = torch.tensor([ok if self.peer_ptr is not None else 0.0], dtype=torch.float64)
        if nranks > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=host_group())
        if flag.item() < 0.5:
            self.peer_ptr = None

    # ------------------------------------------------------------------ set-up
    def _alloc(self, nbytes, count, typestr):
        ptr = ctypes.c_void_p()
        _lib.check(self._lib, self._lib.sb200_peer_alloc(int(nbytes), ctypes.byref(ptr)))
        self._raw.append(ptr.value)
        return torch.as_tensor(_RawDeviceArray(ptr.value, count, typestr), device=self.device)

    def _map_peers(self):
        lib = self._lib
        mine = []
        for t in list(self.local) + [self.flags]:
            handle = ctypes.create_string_buffer(64)
            _lib.check(lib, lib.sb200_ipc_export(ctypes.c_void_p(t.data_ptr()), handle))
            mine.append(handle.raw)
        gathered = [None] * self.nranks
        dist.all_gather_object(gathered, (mine, int(self.device.index or 0)), group=host_group())
        self._dev = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._peer_dev = [g[1] for g in gathered]
        self.peer_ptr, self.peer_flag_ptr = [], []
        for q in range(self.nranks):
            if q == self.rank:
                ptrs = [t.data_ptr() for t in self.local] + [self.flags.data_ptr()]
            else:
                ptrs = []
                for raw in gathered[q][0]:
                    p = ctypes.c_void_p()
                    _lib.check(lib, lib.sb200_ipc_open(ctypes.create_string_buffer(raw, 64), ctypes.byref(p)))
                    self._opened.append(p.value)
                    ptrs.append(p.value)
            self.peer_ptr.append(ptrs[:-1])
            self.peer_flag_ptr.append(ptrs[-1])
        if self.transport == "copy":
            for q in range(self.nranks):
                if q != self.rank:
                    # copy engines: without the reverse mapping the driver stages the copy through the
                    # host (25 GB/s instead of ~640 GB/s, profiles/r01_peer_copy.txt)
                    _lib.check(lib, lib.sb200_enable_peer_access(self._dev, self._peer_dev[q]))
                    _lib.check(lib, lib.sb200_enable_peer_access(self._peer_dev[q], self._dev))

    def close(self, collective=True):
        """Unmap the peers' buffers, then free the own ones.  With ``collective`` (call it on every
        rank) a barrier separates the two, so that no buffer is freed while a peer still maps it."""
        lib = self._lib
        if lib is None or (not self._opened and not self._raw):
            return
        if torch.cuda.is_available():
            torch.cuda.synchronize(self.device)
        for p in self._opened:
            lib.sb200_ipc_close(ctypes.c_void_p(p))
        self._opened = []
        self.peer_ptr = None
        if collective and self.nranks > 1 and dist.is_initialized():
            dist.barrier(group=host_group())
        elif self.nranks > 1:
            # without the barrier a peer may still map these buffers: leave them to process exit rather
            # than free exported memory under an open mapping
            return
        self.local, self.flags = [], None
        for p in self._raw:
            lib.sb200_peer_free(ctypes.c_void_p(p))
        self._raw = []

    def __del__(self):
        try:
            self.close(collective=False)
        except Exception:
            pass

    @property
    def mode(self):
        if self.peer_ptr is None:
            return "nccl all-to-all"
        return "cuda-ipc push kernel" if self.transport != "copy" else "cuda-ipc peer copies"

    # ------------------------------------------------------------------ the exchange
    def exchange(self, dst, src):
        """Block q of this rank's buffer `src` -> block `rank` of rank q's buffer `dst`, for every q
        (buffers are addressed by their index in ``self.local``).  Runs on the current stream."""
        nranks, rank = self.nranks, self.rank
        s_blocks = self.local[src].chunk(nranks)
        if self.peer_ptr is None:
            d_blocks = self.local[dst].chunk(nranks)
            d_blocks[rank].copy_(s_blocks[rank])
            empty = self.local[dst][:0]
            dist.all_to_all([empty if q == rank else d_blocks[q] for q in range(nranks)],
                            [empty if q == rank else s_blocks[q] for q in range(nranks)])
            return
        lib = self._lib
        stream = current_stream_ptr(self.device)
        plan = self._plans.get((dst, src))
        if plan is None:  # pointer tables of this (dst, src) pair, built once
            order = [(rank + k) % nranks for k in range(nranks)]  # local block first, then the ring
            nbytes = s_blocks[0].numel() * 4
            dptr = (ctypes.c_void_p * nranks)(*[self.peer_ptr[q][dst] + rank * nbytes for q in order])
            ddev = (ctypes.c_int * nranks)(*[self._peer_dev[q] for q in order])
            sptr = (ctypes.c_void_p * nranks)(*[s_blocks[q].data_ptr() for q in order])
            # where to signal on each destination: its flags[dst][this rank]
            fptr = (ctypes.c_void_p * nranks)(*[self.peer_flag_ptr[q] + 4 * (dst * nranks + rank) for q in order])
            plan = self._plans[(dst, src)] = (dptr, ddev, sptr, nbytes, fptr)
        dptr, ddev, sptr, nbytes, fptr = plan
        if self.transport == "copy":
            _lib.check(lib, lib.sb200_peer_copy_blocks(nranks, dptr, ddev, sptr, self._dev, nbytes, stream))
        elif self.transport == "push-nccl":
            _lib.check(lib, lib.sb200_peer_push_blocks(nranks, dptr, sptr, nbytes, self.blocks_per_peer, None, 0, None,
                                                       stream))
        else:
            self._epoch[dst] += 1
            epoch = self._epoch[dst]
            _lib.check(lib, lib.sb200_peer_push_blocks(
                nranks, dptr, sptr, nbytes, self.blocks_per_peer, fptr, epoch,
                ctypes.c_void_p(self._done[dst].data_ptr()), stream))
            _lib.check(lib, lib.sb200_peer_wait_flags(
                ctypes.c_void_p(self.flags[dst].data_ptr()), nranks, epoch, ctypes.c_void_p(self._err.data_ptr()),
                stream))
            return
        # all ranks' copies precede their all-reduce in stream order: past this point every block of
        # `dst` has landed here, and every rank has finished reading the `src` blocks it was sent
        dist.all_reduce(self._flag)

    def check(self):
        """raise if a wait ever timed out (synchronises)"""
        if int(self._err.item()) != 0:
            raise _lib.SophtB200Error("peer exchange: a rank did not deliver its blocks in time")
