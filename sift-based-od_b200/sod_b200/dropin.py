"""Bridges between the reference's object-based Python API (lists of cv2.KeyPoint tuples, PoseBin
objects) and the array-based device engine.  Used by the drop-in modules main.py,
HoughTransform.py, HoughTransformHelperFunctions.py and AffineParameters.py.

Only packing/unpacking happens on the host; every computation is a kernel of libsod_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import engine as E
from ._capi import check, lib


def decode_base_bin(packed: int) -> tuple[int, int, int, int]:
    p = int(packed) & 0xFFFFFFFF
    return p & 0xFF, (p >> 8) & 0xFF, (p >> 16) & 0xFF, p >> 24


def tuples_to_scene(matching_keypoints, width: int, height: int) -> E.SceneArrays:
    """[(kpM, kpQ, (w,h), (cx,cy)), ...] -> SceneArrays with one query keypoint, one model keypoint
    and one "image" per tuple (match i <-> row i everywhere).  Only .pt/.angle/.octave are read."""
    n = len(matching_keypoints)
    q_xy = np.empty((n, 2), np.float32)
    m_xy = np.empty((n, 2), np.float32)
    q_ang = np.empty(n, np.float32)
    m_ang = np.empty(n, np.float32)
    q_oct = np.empty(n, np.int32)
    m_oct = np.empty(n, np.int32)
    cent = np.empty((n, 2), np.float64)
    size = np.empty((n, 2), np.float64)
    for i, t in enumerate(matching_keypoints):
        kpm, kpq = t[0], t[1]
        m_xy[i] = kpm.pt
        q_xy[i] = kpq.pt
        m_ang[i] = kpm.angle
        q_ang[i] = kpq.angle
        m_oct[i] = kpm.octave
        q_oct[i] = kpq.octave
        size[i] = t[2]
        cent[i] = t[3]
    return E.SceneArrays(q_xy, q_ang, q_oct, m_xy, m_ang, m_oct, np.arange(n, dtype=np.int32), cent, size,
                         np.array([[width, height]], np.int32))


def estimate_poses(matching_keypoints, width: int = 1, height: int = 1, bins: int = 15):
    """-> (pose float64 [n,4], base bins [(ix,iy,it,is)]) through sod_estimate_pose."""
    n = len(matching_keypoints)
    if n == 0:
        return np.zeros((0, 4)), []
    sc = tuples_to_scene(matching_keypoints, width, height)
    ids = torch.arange(n, dtype=torch.int32, device=sc.device)
    pose = torch.empty((n, 4), dtype=torch.float64, device=sc.device)
    base = torch.empty(n, dtype=torch.int32, device=sc.device)
    lut = torch.tensor(E.sigma_lut(bins), dtype=torch.int32, device=sc.device)
    s = sc.struct()
    check(lib.sod_estimate_pose(C.byref(s), E._ptr(ids), E._ptr(ids), n, bins, E._ptr(lut), E._ptr(pose),
                                E._ptr(base), None, E._stream()), "sod_estimate_pose")
    return pose.cpu().numpy(), [decode_base_bin(b) for b in base.cpu().numpy()]


def pose_bin_indices(poses, bins: int, width: int, height: int):
    p = torch.as_tensor(np.asarray(poses, np.float64).reshape(-1, 4)).cuda()
    n = int(p.shape[0])
    base = torch.empty(max(n, 1), dtype=torch.int32, device=p.device)
    lut = torch.tensor(E.sigma_lut(bins), dtype=torch.int32, device=p.device)
    check(lib.sod_pose_bin_index(E._ptr(p), n, bins, int(width), int(height), E._ptr(lut), E._ptr(base),
                                 E._stream()), "sod_pose_bin_index")
    return [decode_base_bin(b) for b in base[:n].cpu().numpy()]


def hough_dict(matching_keypoints, width: int, height: int, bins, posebin_cls) -> dict:
    """Main.apply_hough_transform on the GPU: returns {(ix,iy,it,is): PoseBin} in the reference's
    insertion order.  Each PoseBin also carries `_sod_members` (indices into matching_keypoints).
    bins: one count, or (bin_x, bin_y, bin_theta, bin_sigma) for the legacy entry point."""
    n = len(matching_keypoints)
    if n == 0:
        return {}
    sc = tuples_to_scene(matching_keypoints, width, height)
    ids = torch.arange(n, dtype=torch.int32, device=sc.device)
    res = E.HoughVoter(sc, bins).vote(ids, ids)
    h = res.host()
    pose = res.pose[:n].cpu().numpy()
    out = {}
    code = h["code"].astype(np.int64)
    _, by, bt, bs = res.dims
    keys = np.stack([code // (by * bt * bs), code // (bt * bs) % by, code // bs % bt, code % bs], 1)
    for i in range(h["n_bins"]):
        cnt = int(h["count"][i])
        off = int(h["offset"][i])
        mem = h["members"][off:off + cnt]
        pairs = [(matching_keypoints[m][0], matching_keypoints[m][1]) for m in mem]
        key = tuple(int(v) for v in keys[i])
        mean = h["mean"][i]
        if cnt == 1:  # the reference keeps the caller's tuple objects until the first update
            first = matching_keypoints[int(mem[0])]
            img_size = first[2]
            mean4 = tuple(float(v) for v in pose[int(mem[0])])
        else:
            img_size = (float(mean[4]), float(mean[5]))
            mean4 = (float(mean[0]), float(mean[1]), float(mean[2]), float(mean[3]))
        pb = posebin_cls(key, img_size, cnt, pairs, mean4)
        pb._sod_members = mem.copy()
        out[key] = pb
    return out


class _PairBatch:
    """Explicit bins (lists of (kpM, kpQ) pairs) packed as a degenerate Hough output so that
    sod_affine_verify can run on PoseBin objects handed in by the caller."""

    def __init__(self, bins_list, width: int, height: int):
        counts = [len(b.keypoint_pairs) for b in bins_list]
        total = int(sum(counts))
        m_xy = np.empty((max(total, 1), 2), np.float32)
        q_xy = np.empty((max(total, 1), 2), np.float32)
        k = 0
        for b in bins_list:
            for pm, pq in b.keypoint_pairs:
                m_xy[k] = pm.pt
                q_xy[k] = pq.pt
                k += 1
        dev = torch.device("cuda")
        z = np.zeros(max(total, 1), np.float32)
        zi = np.zeros(max(total, 1), np.int32)
        self.scene = E.SceneArrays(q_xy, z, zi, m_xy, z, zi, zi, np.zeros((1, 2)), np.zeros((1, 2)),
                                   np.array([[width, height]], np.int32))
        self.total = total
        self.n_bins = len(bins_list)
        off = np.zeros(self.n_bins + 1, np.int64)
        off[1:] = np.cumsum(counts)
        self.offsets = off
        nb = max(self.n_bins, 1)
        self.h = E.HoughResult.__new__(E.HoughResult)
        h = self.h
        h.bins = 1 << 30          # bin_code % bins == bin_code: the code carries pose[3] directly
        h.m_cap = total
        h.cap_bins = nb
        h.cap_votes = max(total, 1)
        h.pose = h.base_bin = h.near_edge = h.bin_order = h.bin_mean = None
        h.counters = torch.tensor([self.n_bins, total, 0, 0, 0, 0, 0, 0], dtype=torch.int32, device=dev)
        h.bin_group = torch.zeros(nb, dtype=torch.int32, device=dev)
        h.bin_code = torch.tensor([int(b.pose[3]) for b in bins_list] or [0], dtype=torch.int32, device=dev)
        h.bin_count = torch.tensor(counts or [0], dtype=torch.int32, device=dev)
        h.bin_offset = torch.tensor(off[:-1].astype(np.int32) if self.n_bins else [0], dtype=torch.int32, device=dev)
        h.members = torch.arange(max(total, 1), dtype=torch.int32, device=dev)
        self.ids = h.members


def affine_run(bins_list, image_query_size, factor_x, factor_y, threshold: int, max_passes: int):
    """Run the fit / prune iteration on explicit PoseBin objects.  Returns per bin (in order):
    (params float64[6] or None, keep mask bool[n_pairs], votes_left, live)."""
    if not bins_list:
        return []
    w, h_img = int(image_query_size[0]), int(image_query_size[1])
    batch = _PairBatch(bins_list, w, h_img)
    if batch.total == 0:
        return [(None, np.zeros(0, bool), 0, 0 >= threshold) for _ in bins_list]
    res = E.affine_verify(batch.scene, batch.ids, batch.ids, batch.h, vote_threshold=0,
                          affine_threshold=threshold, factor=factor_x, factor_y=factor_y,
                          max_passes=max_passes)
    a = res.host(batch.total)
    by_rec = {int(r): i for i, r in enumerate(a["valid_bin"])}
    out = []
    for b in range(batch.n_bins):
        i = by_rec[b]
        sl = slice(int(batch.offsets[b]), int(batch.offsets[b + 1]))
        n_pairs = sl.stop - sl.start
        params = a["params"][i].copy() if n_pairs else None
        out.append((params, a["member_keep"][sl], int(a["votes"][i]), bool(a["live"][i])))
    return out


def residual_keep(model_xy, query_xy, params, x_ref: float, y_ref: float) -> np.ndarray:
    n = len(model_xy)
    if n == 0:
        return np.zeros(0, bool)
    m = torch.as_tensor(np.asarray(model_xy, np.float32).reshape(-1, 2)).cuda()
    q = torch.as_tensor(np.asarray(query_xy, np.float32).reshape(-1, 2)).cuda()
    p = torch.as_tensor(np.asarray(params, np.float64).reshape(6)).cuda()
    keep = torch.empty(n, dtype=torch.uint8, device=m.device)
    check(lib.sod_affine_residual_keep(E._ptr(m), E._ptr(q), n, E._ptr(p), float(x_ref), float(y_ref),
                                       E._ptr(keep), E._stream()), "sod_affine_residual_keep")
    return keep.cpu().numpy().astype(bool)
