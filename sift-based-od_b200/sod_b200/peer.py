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


class PeerBuffer:
    """`nbytes` of zeroed device memory owned by the library (cudaMalloc), exported as a CUDA IPC handle and
    mapped by every other rank of the node: .ptrs[r] is rank r's buffer as this process sees it."""

    def __init__(self, nbytes: int, rank: int, world: int, group=None, device="cuda"):
        import torch.distributed as dist
        self.nbytes, self.rank, self.world, self.group = int(nbytes), int(rank), int(world), group
        self.device = torch.device(device)
        own = C.c_void_p()
        self._own, self._mapped, self.ptrs = None, [], []
        error = None
        with torch.cuda.device(self.device):
            # Every step that can fail on ONE rank only (allocation, export, mapping) is followed by a collective
            # agreement, so that the ranks either all use peer memory or all raise - never a mix that would
            # leave some of them waiting in a collective the others do not enter.
            handle = (C.c_uint8 * HANDLE_BYTES)()
            try:
                check(lib.sod_exchange_alloc(self.nbytes, C.byref(own)), "sod_exchange_alloc")
                self._own = own
                check(lib.sod_ipc_export(own, C.cast(handle, C.c_void_p)), "sod_ipc_export")
            except Exception as exc:
                error = exc
            handles = [None] * self.world
            dist.all_gather_object(handles, None if error else bytes(handle), group=group)
            if error is None and all(h is not None for h in handles):
                try:
                    for r, h in enumerate(handles):
                        if r == self.rank:
                            self.ptrs.append(own.value)
                            continue
                        buf = (C.c_uint8 * HANDLE_BYTES).from_buffer_copy(h)
                        p = C.c_void_p()
                        check(lib.sod_ipc_open(C.cast(buf, C.c_void_p), C.byref(p)), "sod_ipc_open")
                        self._mapped.append(p)
                        self.ptrs.append(p.value)
                except Exception as exc:
                    error = exc
            elif error is None:
                error = RuntimeError("a peer rank could not allocate or export its exchange buffer")
            flags = [None] * self.world
            dist.all_gather_object(flags, error is None, group=group)   # also the barrier: everybody has mapped
            if not all(flags):
                for p in self._mapped:
                    lib.sod_ipc_close(p)
                if self._own is not None:
                    lib.sod_exchange_free(self._own)
                self._own, self._mapped, self.ptrs = None, [], []
                raise RuntimeError(f"peer memory unavailable on rank(s) {[r for r, f in enumerate(flags) if not f]}: "
                                   f"{error!r}")

    def table(self, byte_offset: int = 0):
        """ctypes array of `world` device pointers, every rank's buffer + byte_offset."""
        return (C.c_void_p * self.world)(*[p + int(byte_offset) for p in self.ptrs])

    def own_tensor(self, dtype: torch.dtype, numel: int, byte_offset: int = 0) -> torch.Tensor:
        """A torch view of this rank's own buffer (no copy; the buffer outlives the tensor's users)."""
        item = torch.empty(0, dtype=dtype).element_size()
        typestr = {torch.int32: "<i4", torch.int64: "<i8", torch.uint8: "|u1", torch.float32: "<f4"}[dtype]

        class _Raw:
            __cuda_array_interface__ = {"shape": (int(numel),), "typestr": typestr, "version": 2,
                                        "data": (self.ptrs[self.rank] + int(byte_offset), False)}
        assert byte_offset + numel * item <= self.nbytes
        return torch.as_tensor(_Raw(), device=self.device)

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


class PeerThresholds:
    """The pruning thresholds of a database-sharded run in peer memory (sod_match_top2_peer): two arrays per
    rank, alternating from batch to batch; a sweep min-reduces what it finds into every rank's array."""

    def __init__(self, max_query: int, rank: int, world: int, group=None, device="cuda"):
        self.cap = max(int(lib.sod_row_thr_ints(max_query)), 1)
        self.buf = PeerBuffer(2 * self.cap * 4, rank, world, group, device)
        self.rank, self.world = rank, world
        self._views = [self.buf.own_tensor(torch.int32, self.cap, k * self.cap * 4) for k in range(2)]
        self._tables = {}
        self.batch = 0

    def begin_batch(self) -> torch.Tensor:
        """Next batch: this rank's array of the batch's parity, reset to "none"."""
        self.batch += 1
        thr = self._views[self.batch & 1]
        thr.fill_(0x7F7F7F7F)
        return thr

    def table(self, row_offset: int = 0):
        """Pointers to every rank's array of the current batch, starting at query row `row_offset`."""
        key = (self.batch & 1, int(row_offset))
        t = self._tables.get(key)
        if t is None:
            t = self._tables[key] = self.buf.table(((self.batch & 1) * self.cap + int(row_offset)) * 4)
        return t

    def close(self) -> None:
        self.buf.close()


class PeerExchange:
    def __init__(self, max_query: int, rank: int, world: int, group=None, device="cuda"):
        self.max_query, self.rank, self.world = int(max_query), int(rank), int(world)
        self.device = torch.device(device)
        nbytes = int(lib.sod_exchange_bytes(self.max_query, self.world))
        if nbytes == 0:
            raise ValueError(f"peer exchange supports at most 16 ranks, got {world}")
        self.buf = PeerBuffer(nbytes, rank, world, group, device)
        self._table = self.buf.table()
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
        self.buf.close()
