"""Seeded synthetic inputs for the hot path: descriptor sets, keypoints with planted
object instances, and match lists.  Shared by the tests, tests/golden/make_golden.py and bench.py.
Pure numpy; no reference, oracle or product imports."""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


def sift_like(rng: np.random.Generator, n: int) -> np.ndarray:
    """u8 descriptors with SIFT statistics: |N(0,1)|, L2-normalise, clip 0.2, renormalise, x512."""
    x = np.abs(rng.standard_normal((n, 128)))
    x /= np.linalg.norm(x, axis=1, keepdims=True) + 1e-12
    x = np.minimum(x, 0.2)
    x /= np.linalg.norm(x, axis=1, keepdims=True) + 1e-12
    return np.clip(np.rint(x * 512.0), 0, 255).astype(np.uint8)


def pack_octave(octave: np.ndarray, layer: np.ndarray) -> np.ndarray:
    """cv2 SIFT KeyPoint.octave packing: low byte = octave (two's complement), next byte = layer."""
    return ((octave.astype(np.int64) & 0xFF) | ((layer.astype(np.int64) & 0xFF) << 8)).astype(np.int32)


@dataclass
class SyntheticScene:
    # model database
    m_des: np.ndarray        # u8 [Ndb,128]
    m_xy: np.ndarray         # f32 [Ndb,2]
    m_angle: np.ndarray      # f32 [Ndb] degrees in [0,360)
    m_octave: np.ndarray     # i32 [Ndb] packed
    m_size: np.ndarray       # f32 [Ndb]
    m_image: np.ndarray      # i32 [Ndb] model image of each keypoint
    img_size: np.ndarray     # i32 [n_img,2] (w,h)
    img_centroid: np.ndarray  # f64 [n_img,2]
    # query
    q_des: np.ndarray
    q_xy: np.ndarray
    q_angle: np.ndarray
    q_octave: np.ndarray
    q_size: np.ndarray
    width: int
    height: int
    # ground truth
    true_q: np.ndarray       # query keypoints that are planted matches
    true_t: np.ndarray       # their model keypoints


def make_scene(seed: int, n_images: int = 1, kp_per_image: int = 3000, n_query: int = 800,
               n_true: int = 200, width: int = 2000, height: int = 1500, model_w: int = 1500,
               model_h: int = 1000, scales=(0.5,), noise_px: float = 1.5, desc_noise: int = 3,
               n_dup: int = 0, n_false: int = 0, jitter_frac: float = 0.0,
               jitter_px: float = 40.0) -> SyntheticScene:
    """A query image containing one instance of each of the first len(scales) model images.

    Instance i is model image i mapped by a similarity (scale scales[i] = 2^k, random rotation and
    translation) plus N(0, noise_px) jitter; octaves are chosen so that 2^(q_oct - m_oct) equals the
    scale, angles rotate with the object, descriptors are the model's plus integer noise.  The rest
    of the query keypoints are clutter.  n_dup model rows are duplicated to create exact ties;
    n_false clutter keypoints get the descriptor of a random model row (they pass the ratio test
    but vote at random poses); a fraction jitter_frac of the planted matches is displaced by
    N(0, jitter_px) so that the affine stage has outliers to remove.
    """
    rng = np.random.default_rng(seed)
    ndb = n_images * kp_per_image
    m_des = sift_like(rng, ndb)
    m_xy = np.stack([rng.uniform(0, model_w, ndb), rng.uniform(0, model_h, ndb)], 1).astype(np.float32)
    m_angle = rng.uniform(0, 360, ndb).astype(np.float32)
    m_oct = rng.integers(0, 3, ndb)
    m_layer = rng.integers(1, 4, ndb)
    m_image = np.repeat(np.arange(n_images, dtype=np.int32), kp_per_image)
    img_size = np.tile(np.array([[model_w, model_h]], np.int32), (n_images, 1))
    # GenerateDatabaseInfo.py:33 / SiftHelperFunctions.py:4-14: centroid = mean of keypoint coords,
    # accumulated left to right in Python floats
    cent = np.zeros((n_images, 2), np.float64)
    for i in range(n_images):
        sx = sy = 0.0
        pts = m_xy[i * kp_per_image:(i + 1) * kp_per_image]
        for p in pts:
            sx = sx + float(p[0])
            sy = sy + float(p[1])
        cent[i] = (sx / len(pts), sy / len(pts))
    if n_dup:
        src = rng.integers(0, ndb, n_dup)
        dst = rng.integers(0, ndb, n_dup)
        m_des[dst] = m_des[src]

    q_des = sift_like(rng, n_query)
    q_xy = np.stack([rng.uniform(0, width, n_query), rng.uniform(0, height, n_query)], 1).astype(np.float32)
    q_angle = rng.uniform(0, 360, n_query).astype(np.float32)
    q_oct = rng.integers(-1, 3, n_query)
    q_layer = rng.integers(1, 4, n_query)

    true_q = rng.choice(n_query, n_true, replace=False)
    true_t = np.empty(n_true, np.int64)
    per = n_true // len(scales)
    for i, s in enumerate(scales):
        sl = slice(i * per, (i + 1) * per if i < len(scales) - 1 else n_true)
        k = int(round(math.log2(s)))
        assert 2.0 ** k == s
        t = rng.choice(np.arange(i * kp_per_image, (i + 1) * kp_per_image), sl.stop - sl.start, replace=False)
        true_t[sl] = t
        th = rng.uniform(0, 2 * math.pi)
        c_q = np.array([rng.uniform(0.3, 0.7) * width, rng.uniform(0.3, 0.7) * height])
        rel = (m_xy[t].astype(np.float64) - cent[i]) * s
        rot = np.stack([math.cos(th) * rel[:, 0] - math.sin(th) * rel[:, 1],
                        math.sin(th) * rel[:, 0] + math.cos(th) * rel[:, 1]], 1)
        qi = true_q[sl]
        jit = rng.normal(0, noise_px, rot.shape)
        far = rng.random(len(t)) < jitter_frac
        jit[far] = rng.normal(0, jitter_px, (int(far.sum()), 2))
        q_xy[qi] = (rot + c_q + jit).astype(np.float32)
        q_angle[qi] = np.mod(m_angle[t].astype(np.float64) + math.degrees(th), 360.0).astype(np.float32)
        q_angle[qi] = np.where(q_angle[qi] >= 360.0, 0.0, q_angle[qi])
        q_oct[qi] = m_oct[t] + k          # scale_factor = 2^(q_oct - m_oct) = s
        q_des[qi] = np.clip(m_des[t].astype(np.int16) + rng.integers(-desc_noise, desc_noise + 1, (len(t), 128)),
                            0, 255).astype(np.uint8)
    if n_false:
        free = np.setdiff1d(np.arange(n_query), true_q)
        fq = rng.choice(free, n_false, replace=False)
        ft = rng.integers(0, ndb, n_false)
        q_des[fq] = np.clip(m_des[ft].astype(np.int16) + rng.integers(-desc_noise, desc_noise + 1, (n_false, 128)),
                            0, 255).astype(np.uint8)
    return SyntheticScene(
        m_des=m_des, m_xy=m_xy, m_angle=m_angle, m_octave=pack_octave(m_oct, m_layer),
        m_size=(1.6 * 2.0 ** m_oct).astype(np.float32), m_image=m_image, img_size=img_size,
        img_centroid=cent, q_des=q_des, q_xy=q_xy, q_angle=q_angle,
        q_octave=pack_octave(q_oct, q_layer), q_size=(1.6 * 2.0 ** q_oct).astype(np.float32),
        width=width, height=height, true_q=true_q, true_t=true_t)


def make_match_stress(seed: int, n_objects: int = 500, per_object: int = 4000, inlier_frac: float = 0.1,
                      width: int = 4032, height: int = 3024, model_w: int = 1500, model_h: int = 1000,
                      noise_px: float = 2.0):
    """ratio-passing matches generated directly — per object `per_object` matches, a
    fraction of them consistent with one random similarity (scale in {1,2,4}), the rest uniform.
    Every match has its own query and model keypoint (match i <-> query kp i <-> model kp i).
    Returns a dict of arrays in the layout of oracle.Scene plus match_q/match_t."""
    rng = np.random.default_rng(seed)
    m = n_objects * per_object
    obj = np.repeat(np.arange(n_objects, dtype=np.int32), per_object)
    m_xy = np.stack([rng.uniform(0, model_w, m), rng.uniform(0, model_h, m)], 1).astype(np.float32)
    m_angle = rng.uniform(0, 360, m).astype(np.float32)
    m_oct = rng.integers(0, 3, m)
    q_xy = np.stack([rng.uniform(0, width, m), rng.uniform(0, height, m)], 1).astype(np.float32)
    q_angle = rng.uniform(0, 360, m).astype(np.float32)
    q_oct = rng.integers(-1, 5, m)
    cent = np.stack([m_xy[:, 0].astype(np.float64).reshape(n_objects, per_object).mean(1),
                     m_xy[:, 1].astype(np.float64).reshape(n_objects, per_object).mean(1)], 1)
    n_in = int(per_object * inlier_frac)
    for o in range(n_objects):
        sel = o * per_object + rng.choice(per_object, n_in, replace=False)
        k = int(rng.integers(0, 3))
        s = 2.0 ** k
        th = rng.uniform(0, 2 * math.pi)
        c_q = np.array([rng.uniform(0.2, 0.8) * width, rng.uniform(0.2, 0.8) * height])
        rel = (m_xy[sel].astype(np.float64) - cent[o]) * s
        rot = np.stack([math.cos(th) * rel[:, 0] - math.sin(th) * rel[:, 1],
                        math.sin(th) * rel[:, 0] + math.cos(th) * rel[:, 1]], 1)
        q_xy[sel] = (rot + c_q + rng.normal(0, noise_px, rot.shape)).astype(np.float32)
        a = np.mod(m_angle[sel].astype(np.float64) + math.degrees(th), 360.0).astype(np.float32)
        q_angle[sel] = np.where(a >= 360.0, 0.0, a)
        q_oct[sel] = m_oct[sel] + k
    perm = rng.permutation(m)  # matches arrive in query order, not grouped by object
    one = np.ones(m, np.int64)
    return dict(
        q_xy=q_xy[perm], q_angle=q_angle[perm], q_octave=pack_octave(q_oct, one)[perm],
        m_xy=m_xy, m_angle=m_angle, m_octave=pack_octave(m_oct, one), m_image=obj,
        img_centroid=cent, img_size=np.tile(np.array([[model_w, model_h]], np.int32), (n_objects, 1)),
        img_group=np.arange(n_objects, dtype=np.int32), width=width, height=height,
        match_q=np.arange(m, dtype=np.int32), match_t=perm.astype(np.int32))
