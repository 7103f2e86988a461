"""Host layer over the C ABI: torch owns device memory and streams, every kernel is ours.

Nothing here computes on the CPU; a missing GPU or library raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _capi
from ._capi import check, lib

RATIO = 0.75  # main.py:82


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: torch.Tensor | None) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _require_cuda(t: torch.Tensor, dtype: torch.dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name} must live on the GPU (no CPU path exists)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    return t.contiguous()


def pack_descriptors(des, device: torch.device | str = "cuda") -> torch.Tensor:
    """float32/uint8 [n,128] (numpy or torch, host or device) -> device u8 [n,128].

    OpenCV SIFT descriptors are float32 holding integers 0..255; float input is converted
    on the GPU and rejected if any value is not such an integer.
    """
    t = torch.as_tensor(des)
    if t.ndim != 2 or t.shape[1] != _capi.DESC_DIM:
        raise ValueError(f"descriptors must be [n,{_capi.DESC_DIM}], got {tuple(t.shape)}")
    if t.dtype == torch.uint8:
        return t.to(device, non_blocking=True).contiguous()
    if t.dtype != torch.float32:
        raise TypeError(f"descriptors must be float32 or uint8, got {t.dtype}")
    src = t.to(device, non_blocking=True).contiguous()
    dst = torch.empty(src.shape, dtype=torch.uint8, device=src.device)
    flag = torch.zeros(1, dtype=torch.int32, device=src.device)
    check(lib.sod_pack_u8_from_f32(_ptr(src), src.shape[0], _ptr(dst), _ptr(flag), _stream()),
          "sod_pack_u8_from_f32")
    if int(flag.item()) != 0:
        raise ValueError("descriptors are not integer-valued in 0..255: the exact u8 tensor-core path "
                         "does not apply (bf16 path not built)")
    return dst


@dataclass
class DescriptorShard:
    """A contiguous slice of the model-descriptor database resident in HBM."""
    des: torch.Tensor        # u8 [n,128]
    cq: torch.Tensor         # int32 [padded n] packed |t|^2 (sod_db_prepare)
    index_base: int          # global index of row 0

    @property
    def n(self) -> int:
        return int(self.des.shape[0])


def prepare_db(des_u8: torch.Tensor, index_base: int = 0) -> DescriptorShard:
    des_u8 = _require_cuda(des_u8, torch.uint8, "database descriptors")
    n = int(des_u8.shape[0])
    cq = torch.empty(max(int(lib.sod_cq_ints(n)), 1), dtype=torch.int32, device=des_u8.device)
    check(lib.sod_db_prepare(_ptr(des_u8), n, _ptr(cq), _stream()), "sod_db_prepare")
    return DescriptorShard(des_u8, cq, int(index_base))


class Matcher:
    """2-NN matcher for one database shard; reuses its scratch between calls."""

    def __init__(self, shard: DescriptorShard):
        self.shard = shard
        self._ws: torch.Tensor | None = None
        self._qn: torch.Tensor | None = None

    def _scratch(self, nq: int) -> tuple[torch.Tensor, torch.Tensor]:
        need = int(lib.sod_match_workspace_bytes(nq, self.shard.n))
        dev = self.shard.des.device
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        if self._qn is None or self._qn.numel() < nq:
            self._qn = torch.empty(max(nq, 1), dtype=torch.int32, device=dev)
        return self._ws, self._qn

    def top2(self, q_u8: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        """-> (idx int32 [nq,2] global rows or -1, d2 int32 [nq,2] squared distances, -1 if none)."""
        q_u8 = _require_cuda(q_u8, torch.uint8, "query descriptors")
        nq = int(q_u8.shape[0])
        ws, qn = self._scratch(nq)
        idx = torch.empty((nq, 2), dtype=torch.int32, device=q_u8.device)
        d2 = torch.empty((nq, 2), dtype=torch.int32, device=q_u8.device)
        s = self.shard
        check(lib.sod_query_prepare(_ptr(q_u8), nq, _ptr(qn), _stream()), "sod_query_prepare")
        check(lib.sod_match_top2(_ptr(q_u8), _ptr(qn), nq, _ptr(s.des), _ptr(s.cq), s.n, s.index_base,
                                 _ptr(idx), _ptr(d2), _ptr(ws), ws.numel(), _stream()),
              "sod_match_top2")
        return idx, d2


def merge_top2(parts_idx: torch.Tensor, parts_d2: torch.Tensor, ratio: float = RATIO):
    """[G,nq,2] candidate lists -> (idx [nq,2], d2 [nq,2], dist f32 [nq,2], pass u8 [nq])."""
    parts_idx = _require_cuda(parts_idx, torch.int32, "parts_idx")
    parts_d2 = _require_cuda(parts_d2, torch.int32, "parts_d2")
    if parts_idx.ndim == 2:
        parts_idx, parts_d2 = parts_idx[None], parts_d2[None]
    g, nq = int(parts_idx.shape[0]), int(parts_idx.shape[1])
    dev = parts_idx.device
    idx = torch.empty((nq, 2), dtype=torch.int32, device=dev)
    d2 = torch.empty((nq, 2), dtype=torch.int32, device=dev)
    dist = torch.empty((nq, 2), dtype=torch.float32, device=dev)
    ok = torch.empty(nq, dtype=torch.uint8, device=dev)
    check(lib.sod_top2_merge(_ptr(parts_idx), _ptr(parts_d2), g, nq, _ptr(idx), _ptr(d2), _ptr(dist),
                             _ptr(ok), float(ratio), _stream()), "sod_top2_merge")
    return idx, d2, dist, ok


def knn_match_ratio(q_u8: torch.Tensor, matcher: Matcher, ratio: float = RATIO):
    """Single-shard convenience: knnMatch(k=2) + ratio flags."""
    idx, d2 = matcher.top2(q_u8)
    return merge_top2(idx[None], d2[None], ratio)
