"""Pose clustering after the hot path, drop-in for PostProcessing.py:4-112.

Host-side for now (it runs on the tens-hundreds of bins that survive the affine
stage).  Same neighbour rule, same visiting order and therefore the same clusters and float sums as
the reference, but the depth-first walk is iterative, so large clusters no longer hit Python's
recursion limit.
"""
import math


def dfs(i, seen, graph, out, items):
    """Pre-order depth-first walk from i appending items[...] to out (reference :4-11)."""
    if i in seen:
        return out
    seen.add(i)
    out.append(items[i])
    stack = [iter(graph[i])]
    while stack:
        for nb in stack[-1]:
            if nb not in seen:
                seen.add(nb)
                out.append(items[nb])
                stack.append(iter(graph[nb]))
                break
        else:
            stack.pop()
    return out


def _components(n, graph, items):
    seen = set()
    groups = []
    for i in range(n):
        if i not in seen:
            groups.append(dfs(i, seen, graph, [], items))
    return groups


def group_position(valid_bins):
    """Clusters of bins whose centroids are mutually within a quarter of the (scaled) model size in
    x and y (reference :14-37)."""
    n = len(valid_bins)
    graph = {b: [] for b in range(n)}
    for b in range(n):
        xb, yb = valid_bins[b].centroid
        wb = valid_bins[b].img_size[0] * valid_bins[b].scale / 4
        hb = valid_bins[b].img_size[1] * valid_bins[b].scale / 4
        for a in range(b):
            xa, ya = valid_bins[a].centroid
            dx, dy = abs(xa - xb), abs(ya - yb)
            if dx <= valid_bins[a].img_size[0] * valid_bins[a].scale / 4 and \
                    dy <= valid_bins[a].img_size[1] * valid_bins[a].scale / 4 and dx <= wb and dy <= hb:
                graph[b].append(a)
                graph[a].append(b)
    return _components(n, graph, valid_bins)


def group_orientation(pose_cluster):
    """Inside every position cluster, sub-clusters of bins whose mean angles differ by <= 1 degree
    (reference :39-63)."""
    result = []
    for cluster in pose_cluster:
        n = len(cluster)
        graph = {b: [] for b in range(n)}
        for b in range(n):
            for a in range(b):
                if abs(math.degrees(cluster[a].angle - cluster[b].angle)) <= 1:
                    graph[a].append(b)
                    graph[b].append(a)
        result.append(_components(n, graph, cluster))
    return result


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
