"""Pose clustering after the hot path, drop-in for PostProcessing.py:4-112.

Same functions, arguments and return structures as the reference.  The O(V^2) neighbour tests of
group_position / group_orientation run on the GPU (sod_pose_adjacency / sod_angle_adjacency,
csrc/sod_cluster.cu); the clusters, their order and the order of every float sum are the
reference's (sod_b200/postprocess.py recovers its depth-first visiting order from the adjacency
bit rows), and large clusters no longer hit Python's recursion limit.
"""
import math

from sod_b200 import postprocess as _pp


def dfs(i, set, dict, pose_set, valid_bins):  # noqa: A002 - the reference's parameter names (:4)
    """Pre-order depth-first walk from i appending valid_bins[...] to pose_set (reference :4-11), iterative."""
    seen, graph = set, dict
    if i in seen:
        return pose_set
    seen.add(i)
    pose_set.append(valid_bins[i])
    stack = [iter(graph[i])]
    while stack:
        for nb in stack[-1]:
            if nb not in seen:
                seen.add(nb)
                pose_set.append(valid_bins[nb])
                stack.append(iter(graph[nb]))
                break
        else:
            stack.pop()
    return pose_set


def group_position(valid_bins):
    """Clusters of bins whose centroids are mutually within a quarter of the (scaled) model size in
    x and y (reference :14-37)."""
    clusters, _ = _pp.cluster_positions([b.centroid[0] for b in valid_bins], [b.centroid[1] for b in valid_bins],
                                        [b.scale for b in valid_bins], [b.img_size[0] for b in valid_bins],
                                        [b.img_size[1] for b in valid_bins])
    return [[valid_bins[i] for i in cl] for cl in clusters]


def group_orientation(pose_cluster):
    """Inside every position cluster, sub-clusters of bins whose mean angles differ by <= 1 degree
    (reference :39-63)."""
    pos = _pp.orientation_subclusters([[b.angle for b in cluster] for cluster in pose_cluster])
    return [[[cluster[p] for p in comp] for comp in comps] for cluster, comps in zip(pose_cluster, pos)]


def find_max_orientation(orientation_cluster):
    """Mean angle of the largest orientation sub-cluster, later ones winning ties (reference :65-82)."""
    out = []
    for subclusters in orientation_cluster:
        best, angle = 0, 0
        for sub in subclusters:
            if len(sub) >= best:
                best = len(sub)
                angle = 0
                for b in sub:
                    angle += b.angle
                angle = angle / len(sub)
        out.append(angle)
    return out


def get_final_pose(pose_cluster, final_orientation_list):
    """(mean centroid, orientation, mean scale, image size of the smallest scaled area) per position
    cluster (reference :86-112)."""
    final = []
    for i, cluster in enumerate(pose_cluster):
        sx = sy = ss = 0
        min_area = math.inf
        min_shape = 0, 0
        for b in cluster:
            sx += b.centroid[0]
            sy += b.centroid[1]
            ss += b.scale
            area = (b.img_size[0] * b.scale) * (b.img_size[1] * b.scale)
            if area < min_area:
                min_area = area
                min_shape = b.img_size
        n = len(cluster)
        final.append(((sx / n, sy / n), final_orientation_list[i], ss / n, min_shape))
    return final
