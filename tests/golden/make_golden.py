"""Generate tests/golden/*.npz by running the REAL reference (/root/reference, read-only) on seeded
synthetic scenes.  Run in the authoring container only:

    python tests/golden/make_golden.py

The reference is imported unmodified.  Three workarounds from SURVEY.md §8c: matplotlib is absent
(stub modules), make_kp fails on cv2 >= 4.5.3 and the dataset is absent (Main fields are injected
instead of calling get_query_features).  Keypoints are cv2.KeyPoint objects whose class_id carries
their index so keypoint pairs can be mapped back to indices.

Each fixture stores the inputs and what the reference produced from them:
  knnMatch indices/distances, ratio survivors, the Hough dict in insertion order (key, votes,
  member pairs, running means), the bins surviving apply_affine_parameters with their parameters,
  and final_pose.
"""
import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT / "tests"))

import scenes  # noqa: E402


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib.pyplot"].subplots = lambda *a, **k: (None, None)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    sys.path.insert(0, str(REF))
    import main as refmain  # the reference's main.py
    return refmain


def keypoints(xy, size, angle, octave):
    import cv2
    return [cv2.KeyPoint(float(xy[i, 0]), float(xy[i, 1]), float(size[i]), float(angle[i]), 0.0,
                         int(octave[i]), i) for i in range(len(xy))]


def ragged(lists, dtype=np.int32):
    off = np.zeros(len(lists) + 1, np.int64)
    off[1:] = np.cumsum([len(x) for x in lists])
    flat = np.array([v for x in lists for v in x], dtype).reshape(-1, 2) if lists else np.zeros((0, 2), dtype)
    return off, flat


def run_reference(refmain, sc: scenes.SyntheticScene, bins=15, vote_thr=5, affine_thr=4):
    import cv2
    m = refmain.Main()
    m.kp = keypoints(sc.m_xy, sc.m_size, sc.m_angle, sc.m_octave)
    m.des = sc.m_des.astype(np.float32)
    m.kp_query = keypoints(sc.q_xy, sc.q_size, sc.q_angle, sc.q_octave)
    m.des_query = sc.q_des.astype(np.float32)
    m.img_size_list = [tuple(int(v) for v in sc.img_size[i]) for i in sc.m_image]
    m.img_centroid_list = [tuple(float(v) for v in sc.img_centroid[i]) for i in sc.m_image]
    m.rgb_query = np.zeros((sc.height, sc.width, 3), np.uint8)
    m.image_query_size = (sc.width, sc.height)

    raw = cv2.BFMatcher().knnMatch(m.des_query, m.des, k=2)
    knn_idx = np.array([[a.trainIdx, b.trainIdx] for a, b in raw], np.int32)
    knn_dist = np.array([[a.distance, b.distance] for a, b in raw], np.float32)

    m.run_matcher()
    match_t = np.array([t[0].class_id for t in m.matching_keypoints], np.int32)
    match_q = np.array([t[1].class_id for t in m.matching_keypoints], np.int32)

    m.apply_hough_transform(bins)
    keys = np.array(list(m.hough_transform.keys()), np.int32).reshape(-1, 4)
    hb = list(m.hough_transform.values())
    votes = np.array([b.votes for b in hb], np.int32)
    means = np.array([[b.centroid[0], b.centroid[1], b.angle, b.scale, b.img_size[0], b.img_size[1]]
                      for b in hb], np.float64).reshape(-1, 6)
    mem_off, mem = ragged([[(p[1].class_id, p[0].class_id) for p in b.keypoint_pairs] for b in hb])

    m.get_valid_bins(vote_thr)
    valid_keys = np.array([b.pose for b in m.valid_bins], np.int32).reshape(-1, 4)
    m.apply_affine_parameters(affine_thr)
    live = m.valid_bins
    live_keys = np.array([b.pose for b in live], np.int32).reshape(-1, 4)
    live_votes = np.array([b.votes for b in live], np.int32)
    live_params = np.array([[float(v) for v in b.affine_parameters] for b in live], np.float64).reshape(-1, 6)
    live_off, live_mem = ragged([[(p[1].class_id, p[0].class_id) for p in b.keypoint_pairs] for b in live])
    n_pairs_after = len(m.keypoint_pairs)

    final_pose = np.zeros((0, 6), np.float64)
    try:
        m.post_process()
        final_pose = np.array([[c[0], c[1], o, s, sh[0], sh[1]] for (c, o, s, sh) in m.final_pose],
                              np.float64).reshape(-1, 6)
    except RecursionError:
        pass
    return dict(knn_idx=knn_idx, knn_dist=knn_dist, match_q=match_q, match_t=match_t,
                bin_keys=keys, bin_votes=votes, bin_means=means, bin_mem_off=mem_off, bin_mem=mem,
                valid_keys=valid_keys, live_keys=live_keys, live_votes=live_votes,
                live_params=live_params, live_mem_off=live_off, live_mem=live_mem,
                n_pairs_after=np.int64(n_pairs_after), final_pose=final_pose,
                bins=np.int32(bins), vote_thr=np.int32(vote_thr), affine_thr=np.int32(affine_thr))


SCENES = {
    # one model image, one instance at scale 1/2 (sigma bins 0/1 only: exercises quirk Q1)
    "scene_single": dict(seed=11, n_images=1, kp_per_image=3000, n_query=1500, n_true=220, scales=(0.5,),
                         n_false=500, jitter_frac=0.15),
    # three model images in ONE Hough space (quirk Q7), instances at scales 2, 1, 4; duplicate rows
    "scene_multi": dict(seed=12, n_images=3, kp_per_image=1500, n_query=2000, n_true=420,
                        scales=(2.0, 1.0, 4.0), n_dup=40, width=4032, height=3024, n_false=600,
                        jitter_frac=0.2, jitter_px=60.0),
    # tiny database: a single 128-row tile, ragged query
    "scene_tiny": dict(seed=13, n_images=1, kp_per_image=100, n_query=37, n_true=30, scales=(2.0,)),
}


def main():
    refmain = import_reference()
    import cv2
    for name, kw in SCENES.items():
        sc = scenes.make_scene(**kw)
        out = run_reference(refmain, sc)
        inputs = {f"in_{k}": v for k, v in sc.__dict__.items()}
        np.savez_compressed(HERE / f"{name}.npz", **inputs, **out,
                            versions=np.array([cv2.__version__, np.__version__, sys.version.split()[0]]))
        print(name, "matches", len(out["match_q"]), "bins", len(out["bin_votes"]), "valid",
              len(out["valid_keys"]), "live", len(out["live_keys"]), "final", len(out["final_pose"]))

    # known-answer facts taken from the reference's own libraries
    import math
    from HoughTransformHelperFunctions import calculate_bin_index
    kat = {}
    kat["sigma_k"] = np.arange(-12, 13)
    kat["sigma_bin15"] = np.array([calculate_bin_index((0.0, 0.0, 0.0, 2.0 ** int(k)), 15, (1, 1))[3]
                                   for k in kat["sigma_k"]], np.int32)
    rng = np.random.default_rng(5)
    d = rng.integers(1, 400000, (20000, 2))
    d.sort(1)
    near = np.arange(1, 3000)
    d = np.concatenate([d, np.stack([9 * near * 2, 16 * near * 2], 1), np.stack([9 * near * 2 + 1, 16 * near * 2], 1),
                        np.stack([9 * near * 2 - 1, 16 * near * 2], 1), np.array([[18, 32]])])
    dist = np.sqrt(d.astype(np.float32))
    kat["ratio_d2"] = d.astype(np.int64)
    kat["ratio_pass"] = np.array([float(a) < 0.75 * float(b) for a, b in dist])
    np.savez_compressed(HERE / "kat.npz", **kat)
    print("kat: sigma", kat["sigma_bin15"].tolist(), "ratio cases", len(d),
          "int-test disagreements", int((kat["ratio_pass"] != (16 * d[:, 0] < 9 * d[:, 1])).sum()))


if __name__ == "__main__":
    main()
