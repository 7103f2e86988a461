"""Peer-memory exchange of the shard-local top-2 lists (include/sod.h, "K3 over peer memory").

One exchange buffer per rank, allocated by the library (cudaMalloc), exported as a CUDA IPC handle,
handed to the other ranks of the node through the process group and mapped by each of them.  After
that a merge is two kernels on the caller's stream - no NCCL call, no host round trip - which is
what the 10k-query latency configuration (BASELINE configs[2]) needs; large batches keep the NCCL
scatter form, whose traffic per rank is 2 x 16 B per query row instead of G x 16 B.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import engine as E
from ._capi import check, lib

HANDLE_BYTES = 64


class PeerExchange:
    def __init__(self, max_query: int, rank: int, world: int, group=None, device="cuda"):
        import torch.distributed as dist
        self.max_query, self.rank, self.world = int(max_query), int(rank), int(world)
        self.device = torch.device(device)
        nbytes = int(lib.sod_exchange_bytes(self.max_query, self.world))
        if nbytes == 0:
            raise ValueError(f"peer exchange supports at most 16 ranks, got {world}")
        own = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.sod_exchange_alloc(nbytes, C.byref(own)), "sod_exchange_alloc")
            self._own = own
            handle = (C.c_uint8 * HANDLE_BYTES)()
            check(lib.sod_ipc_export(own, C.cast(handle, C.c_void_p)), "sod_ipc_export")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            self._mapped = []
            ptrs = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    ptrs.append(own.value)
                    continue
                buf = (C.c_uint8 * HANDLE_BYTES).from_buffer_copy(h)
                p = C.c_void_p()
                check(lib.sod_ipc_open(C.cast(buf, C.c_void_p), C.byref(p)), "sod_ipc_open")
                self._mapped.append(p)
                ptrs.append(p.value)
        self._table = (C.c_void_p * self.world)(*ptrs)
        dist.barrier(group=group)        # every rank has mapped every buffer before anybody stores into one
        self.group = group

    def merge(self, idx: torch.Tensor, d2: torch.Tensor, ratio: float = E.RATIO):
        """This rank's lists [nq,2] -> the global (idx, d2, dist, pass) on every rank."""
        idx = E._require_cuda(idx, torch.int32, "idx")
        d2 = E._require_cuda(d2, torch.int32, "d2")
        nq = int(idx.shape[0])
        dev = idx.device
        oi = torch.empty((nq, 2), dtype=torch.int32, device=dev)
        od = torch.empty((nq, 2), dtype=torch.int32, device=dev)
        dist_f = torch.empty((nq, 2), dtype=torch.float32, device=dev)
        ok = torch.empty(nq, dtype=torch.uint8, device=dev)
        check(lib.sod_top2_exchange_peer(E._ptr(idx), E._ptr(d2), nq, self.rank, self.world,
                                         C.cast(self._table, C.c_void_p), self.max_query, E._ptr(oi), E._ptr(od),
                                         E._ptr(dist_f), E._ptr(ok), float(ratio), E._stream()),
              "sod_top2_exchange_peer")
        return oi, od, dist_f, ok

    def close(self) -> None:
        """Unmap the peers' buffers and free the own one (after a barrier: nobody may still store into it)."""
        if self._own is None:
            return
        import torch.distributed as dist
        torch.cuda.synchronize(self.device)
        if dist.is_initialized():
            dist.barrier(group=self.group)
        for p in self._mapped:
            lib.sod_ipc_close(p)
        lib.sod_exchange_free(self._own)
        self._own, self._mapped = None, []
