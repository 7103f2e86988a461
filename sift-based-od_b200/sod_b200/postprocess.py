"""Pose clustering after the affine stage on arrays.

The O(V^2) neighbour tests of group_position / group_orientation (PostProcessing.py:14-63) run on
the GPU (csrc/sod_cluster.cu: bit-matrix adjacency + union-find labels); the reference's depth-first
visiting order - which fixes the order of its float sums, the order of the clusters and the
tie-break of find_max_orientation - is recovered here from the bit rows with integer bitsets
(V steps of V/64-word operations), so results are identical to the reference's and there is no
recursion limit.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from ._capi import check, lib
from .engine import _ptr, _stream


def _graph(kind: str, arrays, segment, n: int, device, limit: float = 1.0):
    words = (n + 31) // 32
    adj = torch.empty((n, words), dtype=torch.int32, device=device)
    label = torch.empty(n, dtype=torch.int32, device=device)
    dev = [torch.as_tensor(np.ascontiguousarray(a, np.float64)).to(device) for a in arrays]
    seg = None if segment is None else torch.as_tensor(np.ascontiguousarray(segment, np.int32)).to(device)
    if kind == "position":
        check(lib.sod_pose_adjacency(_ptr(dev[0]), _ptr(dev[1]), _ptr(dev[2]), _ptr(dev[3]),
                                     _ptr(seg) if seg is not None else 0, n, _ptr(adj), _ptr(label), _stream()),
              "sod_pose_adjacency")
    else:
        check(lib.sod_angle_adjacency(_ptr(dev[0]), _ptr(seg) if seg is not None else 0, n, float(limit),
                                      _ptr(adj), _ptr(label), _stream()), "sod_angle_adjacency")
    return adj.cpu().numpy().view(np.uint32), label.cpu().numpy()


def _row_bitsets(adj: np.ndarray) -> list[int]:
    """uint32 [n][words] -> one Python int per row (bit j = edge to j)."""
    raw = np.ascontiguousarray(adj.astype("<u4"))
    return [int.from_bytes(raw[i].tobytes(), "little") for i in range(raw.shape[0])]


def preorder_components(rows: list[int]) -> list[list[int]]:
    """Connected components in the reference's order: start nodes ascending, pre-order depth-first
    walk that always takes the lowest unvisited neighbour (PostProcessing.py:4-11 over ascending
    neighbour lists)."""
    n = len(rows)
    seen = 0
    comps = []
    for s in range(n):
        if (seen >> s) & 1:
            continue
        seen |= 1 << s
        comp, stack = [s], [s]
        while stack:
            cand = rows[stack[-1]] & ~seen
            if cand:
                nb = (cand & -cand).bit_length() - 1
                seen |= 1 << nb
                comp.append(nb)
                stack.append(nb)
            else:
                stack.pop()
        comps.append(comp)
    return comps


def cluster_positions(cx, cy, scale, img_w, img_h, segment=None, device="cuda"):
    """group_position on arrays -> (clusters as lists of bin indices in visiting order, labels)."""
    n = len(cx)
    if n == 0:
        return [], np.zeros(0, np.int32)
    # the reference's expression img_size[0] * scale / 4 with Python int * float semantics
    reach_x = [img_w[i] * float(scale[i]) / 4 for i in range(n)]
    reach_y = [img_h[i] * float(scale[i]) / 4 for i in range(n)]
    adj, label = _graph("position", (cx, cy, reach_x, reach_y), segment, n, torch.device(device))
    return preorder_components(_row_bitsets(adj)), label


def orientation_subclusters(angles_per_cluster, device="cuda", max_degrees: float = 1.0):
    """group_orientation: for every position cluster (a list of angles in the cluster's visiting
    order) the sub-clusters as lists of POSITIONS in that list, in the reference's order.  All
    clusters go through one launch: angles are laid out back to back with the cluster number as
    segment id, so a cluster's rows and columns are one contiguous range of the bit matrix."""
    sizes = [len(a) for a in angles_per_cluster]
    n = sum(sizes)
    if n == 0:
        return [[] for _ in sizes]
    flat = np.concatenate([np.asarray(a, np.float64) for a in angles_per_cluster if len(a)])
    segment = np.repeat(np.arange(len(sizes), dtype=np.int32), sizes)
    adj, _ = _graph("angle", (flat,), segment, n, torch.device(device), max_degrees)
    rows = _row_bitsets(adj)
    out, off = [], 0
    for m in sizes:
        mask = (1 << m) - 1
        out.append(preorder_components([(rows[off + p] >> off) & mask for p in range(m)]))
        off += m
    return out


def cluster_orientations(angle, clusters, device="cuda", max_degrees: float = 1.0):
    """group_orientation on arrays: per position cluster, sub-clusters as lists of bin indices."""
    pos = orientation_subclusters([[float(angle[i]) for i in cl] for cl in clusters], device, max_degrees)
    return [[[cl[p] for p in comp] for comp in comps] for cl, comps in zip(clusters, pos)]


def post_process_arrays(cx, cy, scale, angle, img_w, img_h, segment=None, device="cuda"):
    """The four reference steps on arrays -> (clusters, sub_clusters, orientations, final) with
    final[c] = ((cx, cy), orientation, scale, (w, h)) exactly as get_final_pose returns it."""
    clusters, label = cluster_positions(cx, cy, scale, img_w, img_h, segment, device)
    subs = cluster_orientations(angle, clusters, device)
    orientations, final = [], []
    for cl, sc in zip(clusters, subs):
        best, ori = 0, 0
        for sub in sc:                      # later sub-clusters win ties (PostProcessing.py:65-82)
            if len(sub) >= best:
                best = len(sub)
                ori = 0
                for i in sub:
                    ori += float(angle[i])
                ori = ori / len(sub)
        orientations.append(ori)
        sx = sy = ss = 0
        min_area, shape = math.inf, (0, 0)
        for i in cl:                        # PostProcessing.py:86-112
            sx += float(cx[i])
            sy += float(cy[i])
            ss += float(scale[i])
            area = (img_w[i] * float(scale[i])) * (img_h[i] * float(scale[i]))
            if area < min_area:
                min_area, shape = area, (img_w[i], img_h[i])
        m = len(cl)
        final.append(((sx / m, sy / m), ori, ss / m, shape))
    return clusters, subs, orientations, final
