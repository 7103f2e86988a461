"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product (sift-based-od_b200/).

A numpy / plain-Python restatement of the reference hot path of torn8to/sift-based-OD, used by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the checker
and the timed CPU baseline.  Every function cites the reference lines it follows
(paths relative to the reference checkout).

Parity pin.  The reference has no tests or golden vectors, and its arithmetic lives
in third-party libraries without pinned versions (OpenCV BFMatcher, numpy.linalg.pinv, libm via
math).  The pins are therefore outputs of the reference itself, run in the authoring container
(opencv-python-headless 4.13.0.92, numpy 2.3.5, CPython 3.12.3) by tests/golden/make_golden.py and
committed under tests/golden/; tests/test_oracle_golden.py holds this file to them.
"""
from __future__ import annotations

import math

import numpy as np

RATIO = 0.75          # main.py:82
DESC_DIM = 128


# --------------------------------------------------------------------------------------------
# Matching: cv2.BFMatcher().knnMatch(des_query, des, k=2) + ratio loop (main.py:68-86)
# --------------------------------------------------------------------------------------------
def knn2(q: np.ndarray, db: np.ndarray, chunk: int = 2048):
    """Two nearest database rows per query row under L2, ascending, ties -> lowest index.

    Restates cv::BFMatcher::knnMatchImpl -> cv::batchDistance(NORM_L2, K=2) as called from
    main.py:70-71.  Descriptors are integer-valued 0..255, so squared distances are
    exact integers (< 2^24, also exact in OpenCV's float32 accumulation).
    Returns (idx int32 [nq,2] with -1 where the database is too small, d2 int64 [nq,2]).
    """
    q = np.ascontiguousarray(q)
    db = np.ascontiguousarray(db)
    nq, n = q.shape[0], db.shape[0]
    idx = np.full((nq, 2), -1, np.int32)
    d2o = np.full((nq, 2), -1, np.int64)
    if nq == 0 or n == 0:
        return idx, d2o
    dbf = db.astype(np.float64)
    tn = (dbf * dbf).sum(1)
    for s in range(0, nq, chunk):
        qf = q[s:s + chunk].astype(np.float64)
        qn = (qf * qf).sum(1)
        d2 = qn[:, None] + tn[None, :] - 2.0 * (qf @ dbf.T)  # exact: all terms < 2^53
        r = np.arange(d2.shape[0])
        i1 = d2.argmin(1)                                      # first minimum = lowest index
        d1 = d2[r, i1].copy()
        idx[s:s + chunk, 0] = i1
        d2o[s:s + chunk, 0] = d1.astype(np.int64)
        if n >= 2:
            d2[r, i1] = np.inf
            i2 = d2.argmin(1)
            idx[s:s + chunk, 1] = i2
            d2o[s:s + chunk, 1] = d2[r, i2].astype(np.int64)
    return idx, d2o


def knn2_float(q: np.ndarray, db: np.ndarray, chunk: int = 1024):
    """knn2 for arbitrary float32 descriptors: same call (main.py:70-71), squared distances as
    float64 sums of the exact float64 differences (cv2 accumulates the same differences in float32).
    The checker for the bf16 path's stated tolerance.  Returns (idx int32 [nq,2], d2 float64 [nq,2],
    inf / -1 where the database is too small)."""
    q = np.ascontiguousarray(q, np.float64)
    db = np.ascontiguousarray(db, np.float64)
    nq, n = q.shape[0], db.shape[0]
    idx = np.full((nq, 2), -1, np.int32)
    d2o = np.full((nq, 2), np.inf, np.float64)
    if nq == 0 or n == 0:
        return idx, d2o
    for s in range(0, nq, chunk):
        qc = q[s:s + chunk]
        d2 = np.empty((qc.shape[0], n))
        for t in range(0, n, 8192):                      # direct differences: no cancellation
            diff = qc[:, None, :] - db[None, t:t + 8192, :]
            d2[:, t:t + 8192] = np.einsum("ijk,ijk->ij", diff, diff)
        r = np.arange(d2.shape[0])
        i1 = d2.argmin(1)
        idx[s:s + chunk, 0] = i1
        d2o[s:s + chunk, 0] = d2[r, i1]
        if n >= 2:
            d2[r, i1] = np.inf
            i2 = d2.argmin(1)
            idx[s:s + chunk, 1] = i2
            d2o[s:s + chunk, 1] = d2[r, i2]
    return idx, d2o


def match_distance(d2: np.ndarray) -> np.ndarray:
    """DMatch.distance: float32 sqrt of the float32 squared distance."""
    return np.sqrt(np.maximum(d2, 0).astype(np.float32))


def ratio_pass(d2: np.ndarray, idx: np.ndarray | None = None, ratio: float = RATIO) -> np.ndarray:
    """`m.distance < 0.75 * n.distance` in Python float (float64) on float32 distances,
    main.py:81-82.  Not equivalent to the integer test 16*d1 < 9*d2."""
    dist = match_distance(d2).astype(np.float64)
    ok = dist[:, 0] < ratio * dist[:, 1]
    if idx is not None:
        ok &= idx[:, 1] >= 0
    return ok


def merge_top2(parts_idx: np.ndarray, parts_d2: np.ndarray):
    """Global top-2 from per-shard top-2 lists by (d2, idx) order — the exchange step of the
    sharded database (no reference counterpart; equals knn2 on the concatenated database)."""
    g, nq, _ = parts_idx.shape
    ci = parts_idx.transpose(1, 0, 2).reshape(nq, g * 2).astype(np.int64)
    cd = parts_d2.transpose(1, 0, 2).reshape(nq, g * 2).astype(np.int64)
    big = np.int64(1) << 40
    key = np.where(ci >= 0, cd * (np.int64(1) << 32) + ci, big * (np.int64(1) << 20))
    order = np.argsort(key, axis=1, kind="stable")[:, :2]
    r = np.arange(nq)[:, None]
    oi = ci[r, order].astype(np.int32)
    od = cd[r, order]
    od[oi < 0] = -1
    return oi, od


# --------------------------------------------------------------------------------------------
# Hough voting: estimate_object_pose / calculate_bin_index / Main.apply_hough_transform / PoseBin
# --------------------------------------------------------------------------------------------
def unpack_octave(packed: int):
    """(octave, scale): low byte sign-extended, scale = 2^-octave exactly
    (SiftHelperFunctions.py:25-40)."""
    o = int(packed) & 0xFF
    if o >= 128:
        o |= -128
    scale = float(1 / (1 << o)) if o >= 0 else float(1 << -o)
    return o, scale


def estimate_pose(m_pt, m_angle, m_octave, q_pt, q_angle, q_octave, m_centroid):
    """Similarity-transform the model centroid into the query image
    (HoughTransformHelperFunctions.py:4-37).  All arithmetic in Python float = IEEE double, libm
    cos/sin, in the reference's order of operations."""
    _, q_scale = unpack_octave(q_octave)
    _, m_scale = unpack_octave(m_octave)
    s = m_scale / q_scale                                     # :22
    tx = (m_centroid[0] - m_pt[0]) * s                        # :25
    ty = (m_centroid[1] - m_pt[1]) * s                        # :26
    a = math.radians(q_angle - m_angle)                       # :28
    a = (a + 2 * math.pi) % (2 * math.pi)                     # :30
    rx = math.cos(a) * tx - math.sin(a) * ty                  # :31
    ry = math.sin(a) * tx + math.cos(a) * ty                  # :32
    return (rx + q_pt[0], ry + q_pt[1], a, s)                 # :34-36


def _bins4(bins):
    return tuple(int(b) for b in bins) if isinstance(bins, (tuple, list)) else (int(bins),) * 4


def bin_index(pose, bins, img_h, img_w):
    """Base bin of a pose (HoughTransformHelperFunctions.py:39-72): x,y truncate toward zero then
    shift by -1 and clamp; theta modulo; log2 scale over 6.5 octaves, clamped.
    bins may be (bin_x, bin_y, bin_theta, bin_sigma): the same expressions with the count of each
    dimension, as the legacy perform_hough_transform writes them (HoughTransform.py:37-56).  The
    reference cannot run that function (it raises on its first vote), so unequal counts
    are parity-unpinned; equal counts are pinned by the golden fixtures."""
    bx, by, bt, bs = _bins4(bins)
    x, y, theta, s = pose
    ix = min(max(0, int((x * bx) / img_w) - 1), bx - 1)              # :49-52
    iy = min(max(0, int((y * by) / img_h) - 1), by - 1)              # :55-58
    it = int((theta * bt / (2 * math.pi)) % bt)                        # :61-63
    n_oct = 4
    isg = int(math.log(s, 2) / (2 * (n_oct - 1) + 0.5) * bs)         # :66-67
    isg = min(max(0, isg), bs - 1)                                     # :69-70
    return ix, iy, it, isg


def sigma_lut(bins: int, kmin: int = -24, kmax: int = 24):
    """i_sigma for scale factors 2^k, k in [kmin,kmax] — the only values the scale ratio of two
    SIFT octaves can take."""
    return [bin_index((0.0, 0.0, 0.0, 2.0 ** k), bins, 1, 1)[3] for k in range(kmin, kmax + 1)]


class Bin:
    """Restates PoseBin (PoseBin.py:6-54): key, running means, votes, members.
    `members` holds match ids (the reference holds (kpM, kpQ) object pairs)."""
    __slots__ = ("group", "pose", "img_size", "votes", "members", "centroid", "angle", "scale", "affine")

    def __init__(self, group, pose, img_size, member, mean):
        self.group = group
        self.pose = pose
        self.img_size = (img_size[0], img_size[1])
        self.votes = 1
        self.members = [member]
        self.centroid = (mean[0], mean[1])
        self.angle = mean[2]
        self.scale = mean[3]
        self.affine = []

    def update(self, pose_val, img_size, member):
        """update_posebin (PoseBin.py:45-51): sequential running means, then append and count."""
        v = self.votes
        self.centroid = ((self.centroid[0] * v + pose_val[0]) / (v + 1),
                         (self.centroid[1] * v + pose_val[1]) / (v + 1))      # :19-25
        self.angle = (self.angle * v + pose_val[2]) / (v + 1)                # :27-30
        self.scale = (self.scale * v + pose_val[3]) / (v + 1)                # :32-35
        self.img_size = ((self.img_size[0] * v + img_size[0]) / (v + 1),
                         (self.img_size[1] * v + img_size[1]) / (v + 1))      # :37-43
        self.members.append(member)
        self.votes += 1


class Scene:
    """Structure-of-arrays inputs of the Hough/affine stages (what Main holds as object lists).

    q_*: query keypoints; m_*: model keypoints; m_image: model image of each model keypoint;
    img_centroid/img_size: per model image (GenerateDatabaseInfo.py:33-34); img_group: Hough space
    of each image (None = one space for everything, the reference's behaviour);
    width/height of the query image (main.py:45).
    """

    def __init__(self, q_xy, q_angle, q_octave, m_xy, m_angle, m_octave, m_image, img_centroid,
                 img_size, width, height, img_group=None):
        self.q_xy = np.asarray(q_xy, np.float32)
        self.q_angle = np.asarray(q_angle, np.float32)
        self.q_octave = np.asarray(q_octave, np.int32)
        self.m_xy = np.asarray(m_xy, np.float32)
        self.m_angle = np.asarray(m_angle, np.float32)
        self.m_octave = np.asarray(m_octave, np.int32)
        self.m_image = np.asarray(m_image, np.int32)
        self.img_centroid = np.asarray(img_centroid, np.float64)
        self.img_size = np.asarray(img_size, np.float64)
        self.width = int(width)
        self.height = int(height)
        self.img_group = None if img_group is None else np.asarray(img_group, np.int32)

    def pose_of(self, qi: int, ti: int):
        img = int(self.m_image[ti])
        return estimate_pose((float(self.m_xy[ti, 0]), float(self.m_xy[ti, 1])), float(self.m_angle[ti]),
                             int(self.m_octave[ti]), (float(self.q_xy[qi, 0]), float(self.q_xy[qi, 1])),
                             float(self.q_angle[qi]), int(self.q_octave[qi]),
                             (float(self.img_centroid[img, 0]), float(self.img_centroid[img, 1])))


def hough_vote(scene: Scene, match_q, match_t, bins=15):
    """Main.apply_hough_transform (main.py:89-119): every match votes into the 2x2x2x2 bins starting
    at its base bin; candidates with any coordinate >= bins are dropped (no theta wrap).
    Returns an insertion-ordered dict (group, ix, iy, it, is) -> Bin."""
    table: dict = {}
    b4 = _bins4(bins)
    for mid, (qi, ti) in enumerate(zip(match_q, match_t)):
        qi, ti = int(qi), int(ti)
        img = int(scene.m_image[ti])
        group = 0 if scene.img_group is None else int(scene.img_group[img])
        size = (float(scene.img_size[img, 0]), float(scene.img_size[img, 1]))
        pose_val = scene.pose_of(qi, ti)
        ix, iy, it, isg = bin_index(pose_val, bins, scene.height, scene.width)
        for w in range(2):
            for x in range(2):
                for y in range(2):
                    for z in range(2):
                        p = (ix + w, iy + x, it + y, isg + z)
                        if p[0] < b4[0] and p[1] < b4[1] and p[2] < b4[2] and p[3] < b4[3]:   # :110
                            key = (group,) + p
                            b = table.get(key)
                            if b is None:
                                table[key] = Bin(group, p, size, mid, pose_val)             # :119
                            else:
                                b.update(pose_val, size, mid)                               # :113
    return table


def hough_base_bins_vectorized(scene: Scene, match_q, match_t, bins: int = 15):
    """numpy fp64 version of estimate_pose + bin_index for large match sets (2M-match config).
    Same operation order; numpy's cos/sin are not guaranteed bit-identical to libm, so tests use it
    only together with a scalar cross-check.  Returns (pose [M,4], base [M,4] int32)."""
    mq = np.asarray(match_q, np.int64)
    mt = np.asarray(match_t, np.int64)
    img = scene.m_image[mt]
    qo = (scene.q_octave[mq] & 0xFF).astype(np.int64)
    qo = np.where(qo >= 128, qo - 256, qo)
    mo = (scene.m_octave[mt] & 0xFF).astype(np.int64)
    mo = np.where(mo >= 128, mo - 256, mo)
    s = np.exp2((qo - mo).astype(np.float64))
    tx = (scene.img_centroid[img, 0] - scene.m_xy[mt, 0].astype(np.float64)) * s
    ty = (scene.img_centroid[img, 1] - scene.m_xy[mt, 1].astype(np.float64)) * s
    a = (scene.q_angle[mq].astype(np.float64) - scene.m_angle[mt].astype(np.float64)) * (math.pi / 180.0)
    a = np.mod(a + 2 * math.pi, 2 * math.pi)
    ca, sa = np.cos(a), np.sin(a)
    x = (ca * tx - sa * ty) + scene.q_xy[mq, 0].astype(np.float64)
    y = (sa * tx + ca * ty) + scene.q_xy[mq, 1].astype(np.float64)
    ix = np.clip(np.trunc((x * bins) / scene.width).astype(np.int64) - 1, 0, bins - 1)
    iy = np.clip(np.trunc((y * bins) / scene.height).astype(np.int64) - 1, 0, bins - 1)
    it = np.trunc(np.mod(a * bins / (2 * math.pi), bins)).astype(np.int64)
    lut = np.asarray(sigma_lut(bins), np.int64)
    isg = lut[np.clip(qo - mo, -24, 24) + 24]
    return np.stack([x, y, a, s], 1), np.stack([ix, iy, it, isg], 1).astype(np.int32)


def vote_counts_vectorized(base: np.ndarray, group: np.ndarray, bins: int = 15):
    """Bin -> votes from base bins: 16 offsets, `< bins` filter (main.py:105-110).
    Returns (keys int64 sorted, counts) with key = (((g*bins+ix)*bins+iy)*bins+it)*bins+is."""
    keys = []
    for w in range(2):
        for x in range(2):
            for y in range(2):
                for z in range(2):
                    p = base + np.array([w, x, y, z], np.int32)
                    ok = (p < bins).all(1)
                    k = group[ok].astype(np.int64)
                    for d in range(4):
                        k = k * bins + p[ok, d]
                    keys.append(k)
    return np.unique(np.concatenate(keys), return_counts=True)


def valid_bins(table: dict, threshold: int = 5):
    """Main.get_valid_bins (main.py:121-132): bins with votes >= threshold in insertion order."""
    return [b for b in table.values() if b.votes >= threshold]


# --------------------------------------------------------------------------------------------
# Affine verification: AffineParameters / remove_outliers / Main.apply_affine_parameters
# --------------------------------------------------------------------------------------------
def affine_fit(model_xy, query_xy):
    """Least-squares [m1,m2,m3,m4,tx,ty] with u = m1 x + m2 y + tx, v = m3 x + m4 y + ty
    (AffineParameters.py:11-55,89-110): rows [x,y,0,0,1,0] / [0,0,x,y,0,1], x = pinv(A^T A) A^T b,
    built from Python lists exactly as the reference does so BLAS sees the same operands."""
    rows, rhs = [], []
    for (x, y), (u, v) in zip(model_xy, query_xy):
        rows.append([x, y, 0, 0, 1, 0])
        rows.append([0, 0, x, y, 0, 1])
        rhs.append(u)
        rhs.append(v)
    at = np.transpose(rows).tolist()
    g = np.linalg.pinv(np.matmul(at, rows))
    return np.matmul(g, np.matmul(at, rhs))


def affine_residual_keep(params, model_xy, query_xy, x_ref, y_ref):
    """remove_outliers decision per pair (AffineParameters.py:128-155): keep unless
    |u_AP-u| > x_ref or |v_AP-v| > y_ref."""
    m = [[params[0], params[1]], [params[2], params[3]]]
    t = [params[4], params[5]]
    keep = []
    for (x, y), (u, v) in zip(model_xy, query_xy):
        uv = (np.matmul(m, [x, y]) + t).tolist()
        keep.append(not (abs(uv[0] - u) > x_ref or abs(uv[1] - v) > y_ref))
    return keep


def affine_verify(scene: Scene, match_q, match_t, bins_list, threshold: int = 4, pos_factor: int = 32):
    """Main.apply_affine_parameters (main.py:139-157): global fit / prune loop until a full pass
    removes nothing; after each pass only bins with votes >= threshold stay.  The residual limits
    use pose[3], the sigma BIN INDEX: x_ref = W*pose[3]/(4*pos_factor).
    Mutates and returns the surviving Bin list (members pruned, .affine set, means untouched Q6)."""
    factor = pos_factor * 4
    live = list(bins_list)
    changed = True
    while changed:
        changed = False
        nxt = []
        for b in live:
            mxy = [(float(scene.m_xy[int(match_t[i]), 0]), float(scene.m_xy[int(match_t[i]), 1])) for i in b.members]
            qxy = [(float(scene.q_xy[int(match_q[i]), 0]), float(scene.q_xy[int(match_q[i]), 1])) for i in b.members]
            if mxy:
                b.affine = [float(v) for v in affine_fit(mxy, qxy)]
            x_ref = scene.width * b.pose[3] / factor
            y_ref = scene.height * b.pose[3] / factor
            keep = affine_residual_keep(b.affine, mxy, qxy, x_ref, y_ref) if mxy else []
            if not all(keep):
                changed = True
                b.members = [i for i, k in zip(b.members, keep) if k]
            b.votes = len(b.members)
            if b.votes >= threshold:
                nxt.append(b)
        live = nxt
    return live


# --------------------------------------------------------------------------------------------------
# Post-processing: PostProcessing.py:4-112 on arrays
# --------------------------------------------------------------------------------------------------
def _preorder_components(n: int, nbrs) -> list:
    """Components of an undirected graph in the reference's visiting order: start nodes ascending,
    recursive pre-order dfs (PostProcessing.py:4-11) over neighbour lists that are ascending by
    construction (:17-30); written iteratively so large clusters do not hit the recursion limit."""
    seen = [False] * n
    comps = []
    for s in range(n):
        if seen[s]:
            continue
        seen[s] = True
        comp = [s]
        stack = [(s, 0)]
        while stack:
            v, k = stack.pop()
            lst = nbrs[v]
            while k < len(lst) and seen[lst[k]]:
                k += 1
            if k < len(lst):
                nb = lst[k]
                stack.append((v, k + 1))
                seen[nb] = True
                comp.append(nb)
                stack.append((nb, 0))
        comps.append(comp)
    return comps


def post_process(cx, cy, scale, angle, img_w, img_h):
    """group_position -> group_orientation -> find_max_orientation -> get_final_pose on n surviving
    bins given as arrays (centroid, scale, angle = PoseBin running means; img_w/img_h = model image
    size).  Returns (clusters, sub_clusters, orientations, final) with bins as indices:
    clusters[c] = bin indices in visiting order; sub_clusters[c] = lists of bin indices;
    final[c] = (cx, cy, orientation, scale, w, h)."""
    n = len(cx)
    cx = [float(v) for v in cx]; cy = [float(v) for v in cy]
    scale = [float(v) for v in scale]; angle = [float(v) for v in angle]
    # Python ints stay ints in `img_size[0] * scale / 4` (int * float), as in the reference
    nbrs = [[] for _ in range(n)]
    for b in range(n):                                       # PostProcessing.py:17-30
        for a in range(b):
            dx, dy = abs(cx[a] - cx[b]), abs(cy[a] - cy[b])
            if dx <= img_w[a] * scale[a] / 4 and dy <= img_h[a] * scale[a] / 4 and \
                    dx <= img_w[b] * scale[b] / 4 and dy <= img_h[b] * scale[b] / 4:
                nbrs[b].append(a)
                nbrs[a].append(b)
    clusters = _preorder_components(n, nbrs)
    sub_clusters, orientations, final = [], [], []
    for cl in clusters:
        m = len(cl)
        g = [[] for _ in range(m)]
        for b in range(m):                                   # :43-56, positions within the cluster list
            for a in range(b):
                if abs(math.degrees(angle[cl[a]] - angle[cl[b]])) <= 1:
                    g[a].append(b)
                    g[b].append(a)
        subs = [[cl[p] for p in comp] for comp in _preorder_components(m, g)]
        sub_clusters.append(subs)
        best, ori = 0, 0                                     # :65-82 (later sub-clusters win ties)
        for sub in subs:
            if len(sub) >= best:
                best = len(sub)
                ori = 0
                for i in sub:
                    ori += angle[i]
                ori = ori / len(sub)
        orientations.append(ori)
        sx = sy = ss = 0                                     # :86-112
        min_area, shape = math.inf, (0, 0)
        for i in cl:
            sx += cx[i]
            sy += cy[i]
            ss += scale[i]
            area = (img_w[i] * scale[i]) * (img_h[i] * scale[i])
            if area < min_area:
                min_area, shape = area, (img_w[i], img_h[i])
        final.append((sx / m, sy / m, ori, ss / m, shape[0], shape[1]))
    return clusters, sub_clusters, orientations, final


def pipeline(scene: Scene, q_des, db_des, bins: int = 15, vote_threshold: int = 5,
             affine_threshold: int = 4):
    """main.py:180-185 end to end on arrays: match -> ratio -> vote -> valid bins -> affine."""
    idx, d2 = knn2(q_des, db_des)
    ok = ratio_pass(d2, idx)
    mq = np.nonzero(ok)[0].astype(np.int32)
    mt = idx[ok, 0].astype(np.int32)
    table = hough_vote(scene, mq, mt, bins)
    vb = valid_bins(table, vote_threshold)
    # keep the pre-affine state for comparisons
    pre = {(b.group,) + b.pose: (b.votes, list(b.members)) for b in table.values()}
    live = affine_verify(scene, mq, mt, vb, affine_threshold)
    return dict(idx=idx, d2=d2, ok=ok, match_q=mq, match_t=mt, table=table, pre=pre, live=live)
