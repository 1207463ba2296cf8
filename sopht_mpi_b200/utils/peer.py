"""All-to-all of equal blocks over NVLink peer memory (one process per GPU, one node).

Every rank allocates its exchange buffers, publishes them through CUDA IPC and maps the buffers of
all other ranks; an exchange is then ONE small kernel (``sb200_peer_push_blocks``) whose thread blocks
store every destination block straight into the destination rank's buffer, all peers concurrently
through the NVSwitch, followed by a tiny NCCL all-reduce that orders "all blocks have landed" on
every rank's stream.  Replaces the MPI ``Alltoallw`` inside mpi4py-fft's transposes (reference
``poisson_solver_3d/fft_mpi_3d.py:27-48``).

"All blocks have landed" is signalled through peer memory too: the last thread block of every
destination raises an epoch flag in the destination's memory, and a one-warp kernel on the
destination's stream waits for the P flags (bounded spin), so no collective sits on the critical path.

``SB200_EXCHANGE`` selects the transport for measurements: ``push`` (default), ``push-nccl`` (push kernel +
a one-element NCCL all-reduce as the barrier), ``copy`` (one ``cudaMemcpyPeerAsync`` per destination on
the stream + all-reduce, round 1), ``nccl`` (all-to-all).  If CUDA IPC is
not available (different nodes, no peer access) the exchange falls back to NCCL.
"""
import os
import ctypes

import torch
import torch.distributed as dist
from torch.multiprocessing.reductions import reduce_tensor

from .. import _lib
from .comm import host_group
from .device import current_stream_ptr
from .logger import logger


class PeerExchange:
    def __init__(self, n_buffers, n_float32, device, rank, nranks, use_peer_copies=True):
        self.rank, self.nranks, self.device = rank, nranks, device
        # zero-initialised (the padding bins of the last kx block are never written)
        self.local = [torch.zeros(n_float32, dtype=torch.float32, device=device) for _ in range(n_buffers)]
        self.peer = None
        self._plans = {}
        self._flag = torch.zeros(1, dtype=torch.float32, device=device)
        # epoch flags [buffer][source rank] (written by the peers), arrival counters of the push kernel
        self.flags = torch.zeros((n_buffers, max(nranks, 1)), dtype=torch.int32, device=device)
        self._done = torch.zeros((n_buffers, 8), dtype=torch.int32, device=device)
        self._err = torch.zeros(1, dtype=torch.int32, device=device)
        self._epoch = [0] * n_buffers
        self.transport = os.environ.get("SB200_EXCHANGE", "push")
        self.blocks_per_peer = int(os.environ.get("SB200_PUSH_BLOCKS", "0"))
        if self.transport == "nccl":
            use_peer_copies = False
        if use_peer_copies and nranks > 1:
            try:
                self._map_peers()
            except Exception as exc:  # pragma: no cover - depends on the node
                logger.warning(f"CUDA IPC peer mapping unavailable ({type(exc).__name__}: {exc}); "
                               "the transposes use NCCL send/recv")
                self.peer = None
        ok = torch.tensor([1.0 if self.peer is not None else 0.0], dtype=torch.float64)
        if nranks > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=host_group())
        if ok.item() < 0.5:
            self.peer = None

    def _map_peers(self):
        handles = [reduce_tensor(t) for t in self.local] + [reduce_tensor(self.flags)]
        gathered = [None] * self.nranks
        dist.all_gather_object(gathered, handles, group=host_group())
        self.peer, self.peer_flags = [], []
        for q in range(self.nranks):
            if q == self.rank:
                self.peer.append(self.local)
                self.peer_flags.append(self.flags)
            else:
                rebuilt = [fn(*args) for fn, args in gathered[q]]
                self.peer.append(rebuilt[:-1])
                self.peer_flags.append(rebuilt[-1])
        self._keep = gathered  # the rebuilt storages reference the senders' handles
        self._lib = _lib.load()
        self._dev = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._peer_dev = [self.peer[q][0].device.index for q in range(self.nranks)]
        for q in range(self.nranks):
            if q != self.rank:
                # both directions (torch maps the imported buffers in a context on the exporting device,
                # so that context exists in this process anyway): without the reverse mapping the driver
                # stages copies through the host (25 GB/s instead of ~640 GB/s,
                # profiles/r01_peer_copy.txt)
                _lib.check(self._lib, self._lib.sb200_enable_peer_access(self._dev, self._peer_dev[q]))
                _lib.check(self._lib, self._lib.sb200_enable_peer_access(self._peer_dev[q], self._dev))

    @property
    def mode(self):
        if self.peer is None:
            return "nccl all-to-all"
        return "cuda-ipc push kernel" if self.transport != "copy" else "cuda-ipc peer copies"

    def exchange(self, dst, src):
        """Block q of this rank's buffer `src` -> block `rank` of rank q's buffer `dst`, for every q
        (buffers are addressed by their index in ``self.local``).  Runs on the current stream."""
        nranks, rank = self.nranks, self.rank
        s_blocks = self.local[src].chunk(nranks)
        if self.peer is None:
            d_blocks = self.local[dst].chunk(nranks)
            d_blocks[rank].copy_(s_blocks[rank])
            empty = self.local[dst][:0]
            dist.all_to_all([empty if q == rank else d_blocks[q] for q in range(nranks)],
                            [empty if q == rank else s_blocks[q] for q in range(nranks)])
            return
        stream = current_stream_ptr(self.device)
        plan = self._plans.get((dst, src))
        if plan is None:  # pointer tables of this (dst, src) pair, built once
            order = [(rank + k) % nranks for k in range(nranks)]  # local block first, then the ring
            dptr = (ctypes.c_void_p * nranks)(*[self.peer[q][dst].chunk(nranks)[rank].data_ptr() for q in order])
            ddev = (ctypes.c_int * nranks)(*[self._peer_dev[q] for q in order])
            sptr = (ctypes.c_void_p * nranks)(*[s_blocks[q].data_ptr() for q in order])
            # where to signal on each destination: its flags[dst][this rank]
            fptr = (ctypes.c_void_p * nranks)(*[self.peer_flags[q][dst, rank:rank + 1].data_ptr() for q in order])
            plan = self._plans[(dst, src)] = (dptr, ddev, sptr, s_blocks[0].numel() * 4, fptr)
        dptr, ddev, sptr, nbytes, fptr = plan
        if self.transport == "copy":
            _lib.check(self._lib, self._lib.sb200_peer_copy_blocks(nranks, dptr, ddev, sptr, self._dev, nbytes, stream))
        elif self.transport == "push-nccl":
            _lib.check(self._lib, self._lib.sb200_peer_push_blocks(nranks, dptr, sptr, nbytes, self.blocks_per_peer,
                                                                  None, 0, None, stream))
        else:
            self._epoch[dst] += 1
            epoch = self._epoch[dst]
            _lib.check(self._lib, self._lib.sb200_peer_push_blocks(
                nranks, dptr, sptr, nbytes, self.blocks_per_peer, fptr, epoch,
                ctypes.c_void_p(self._done[dst].data_ptr()), stream))
            _lib.check(self._lib, self._lib.sb200_peer_wait_flags(
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
