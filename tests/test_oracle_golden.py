"""The oracle (oracle/sod_oracle.py) held to outputs of the real reference (tests/golden/*.npz,
made by tests/golden/make_golden.py from /root/reference).  CPU only."""
from pathlib import Path

import numpy as np
import pytest

from oracle import sod_oracle as O

GOLD = Path(__file__).resolve().parent / "golden"
SCENES = ["scene_single", "scene_multi", "scene_tiny", "scene_c1"]


def load(name):
    z = np.load(GOLD / f"{name}.npz")
    scene = O.Scene(z["in_q_xy"], z["in_q_angle"], z["in_q_octave"], z["in_m_xy"], z["in_m_angle"],
                    z["in_m_octave"], z["in_m_image"], z["in_img_centroid"], z["in_img_size"],
                    int(z["in_width"]), int(z["in_height"]))
    return z, scene


@pytest.mark.parametrize("name", SCENES)
def test_matching_equals_cv2_knnmatch(name):
    z, _ = load(name)
    idx, d2 = O.knn2(z["in_q_des"], z["in_m_des"])
    np.testing.assert_array_equal(idx, z["knn_idx"])
    np.testing.assert_array_equal(O.match_distance(d2), z["knn_dist"])       # float32, bit-exact
    ok = O.ratio_pass(d2, idx)
    np.testing.assert_array_equal(np.nonzero(ok)[0], z["match_q"])
    np.testing.assert_array_equal(idx[ok, 0], z["match_t"])


@pytest.mark.parametrize("name", SCENES)
def test_hough_dict_equals_reference(name):
    z, scene = load(name)
    table = O.hough_vote(scene, z["match_q"], z["match_t"], int(z["bins"]))
    keys = np.array([k[1:] for k in table.keys()], np.int32).reshape(-1, 4)
    np.testing.assert_array_equal(keys, z["bin_keys"])                       # same insertion order
    bins = list(table.values())
    np.testing.assert_array_equal([b.votes for b in bins], z["bin_votes"])
    means = np.array([[b.centroid[0], b.centroid[1], b.angle, b.scale, b.img_size[0], b.img_size[1]]
                      for b in bins])
    np.testing.assert_array_equal(means, z["bin_means"])                     # float64, bit-exact
    off = z["bin_mem_off"]
    for i, b in enumerate(bins):
        pairs = z["bin_mem"][off[i]:off[i + 1]]
        np.testing.assert_array_equal(z["match_q"][b.members], pairs[:, 0])
        np.testing.assert_array_equal(z["match_t"][b.members], pairs[:, 1])


@pytest.mark.parametrize("name", SCENES)
def test_vectorized_bins_equal_reference(name):
    z, scene = load(name)
    _, base = O.hough_base_bins_vectorized(scene, z["match_q"], z["match_t"], int(z["bins"]))
    keys, counts = O.vote_counts_vectorized(base, np.zeros(len(base), np.int32), int(z["bins"]))
    b = int(z["bins"])
    gk = ((z["bin_keys"][:, 0].astype(np.int64) * b + z["bin_keys"][:, 1]) * b + z["bin_keys"][:, 2]) * b + z["bin_keys"][:, 3]
    order = np.argsort(gk)
    np.testing.assert_array_equal(keys, gk[order])
    np.testing.assert_array_equal(counts, z["bin_votes"][order])


@pytest.mark.parametrize("name", SCENES)
def test_affine_verification_equals_reference(name):
    z, scene = load(name)
    table = O.hough_vote(scene, z["match_q"], z["match_t"], int(z["bins"]))
    vb = O.valid_bins(table, int(z["vote_thr"]))
    np.testing.assert_array_equal(np.array([b.pose for b in vb], np.int32).reshape(-1, 4), z["valid_keys"])
    live = O.affine_verify(scene, z["match_q"], z["match_t"], vb, int(z["affine_thr"]))
    np.testing.assert_array_equal(np.array([b.pose for b in live], np.int32).reshape(-1, 4), z["live_keys"])
    np.testing.assert_array_equal([b.votes for b in live], z["live_votes"])
    np.testing.assert_array_equal(np.array([b.affine for b in live]).reshape(-1, 6), z["live_params"])
    off = z["live_mem_off"]
    for i, b in enumerate(live):
        pairs = z["live_mem"][off[i]:off[i + 1]]
        np.testing.assert_array_equal(z["match_q"][b.members], pairs[:, 0])
        np.testing.assert_array_equal(z["match_t"][b.members], pairs[:, 1])
    assert sum(b.votes for b in live) == int(z["n_pairs_after"])


def test_known_answers():
    z = np.load(GOLD / "kat.npz")
    lut = O.sigma_lut(15, int(z["sigma_k"][0]), int(z["sigma_k"][-1]))
    np.testing.assert_array_equal(lut, z["sigma_bin15"])
    np.testing.assert_array_equal(O.ratio_pass(z["ratio_d2"]), z["ratio_pass"])
    assert O.ratio_pass(np.array([[18, 32]]))[0]


def test_ties_lowest_index_first():
    rng = np.random.default_rng(3)
    db = rng.integers(0, 256, (50, 128), dtype=np.uint8)
    q = db[[7]].copy()
    db[30] = db[40] = db[7]
    idx, d2 = O.knn2(q, db)
    assert idx.tolist() == [[7, 30]] and d2.tolist() == [[0, 0]]


def test_degenerate_sizes():
    q = np.zeros((3, 128), np.uint8)
    idx, d2 = O.knn2(q, np.zeros((0, 128), np.uint8))
    assert (idx == -1).all()
    idx, d2 = O.knn2(q, np.ones((1, 128), np.uint8))
    assert idx.tolist() == [[0, -1]] * 3 and d2[:, 0].tolist() == [128] * 3
    assert not O.ratio_pass(d2, idx).any()


def test_post_process_equals_reference():
    """oracle.post_process vs the reference's PostProcessing functions (fixture postprocess.npz):
    cluster membership in visiting order, orientation sub-clusters, orientations and final poses,
    all bit-exact."""
    z = np.load(GOLD / "postprocess.npz")
    clusters, subs, ori, final = O.post_process(z["cx"], z["cy"], z["scale"], z["angle"],
                                                [int(v) for v in z["img_w"]], [int(v) for v in z["img_h"]])
    off = z["cluster_off"]
    assert [len(c) for c in clusters] == np.diff(off).tolist()
    assert [i for c in clusters for i in c] == z["cluster_members"].tolist()
    assert [len(s) for s in subs] == z["subs_per_cluster"].tolist()
    flat = [sub for s in subs for sub in s]
    assert [len(sub) for sub in flat] == np.diff(z["sub_off"]).tolist()
    assert [i for sub in flat for i in sub] == z["sub_members"].tolist()
    np.testing.assert_array_equal(np.array(ori, np.float64), z["orientation"])
    np.testing.assert_array_equal(np.array(final, np.float64), z["final"])


def test_affine_on_rank_deficient_bins_equals_reference():
    """tests/golden/affine_singular.npz: the reference's AffineParameters + the fixed-point loop of
    Main.apply_affine_parameters on bins with duplicate model locations, collinear points, threshold
    sizes and ill-conditioned normal matrices.  The oracle restates them bit for bit."""
    z = np.load(GOLD / "affine_singular.npz")
    off = z["off"]
    n = len(z["names"])
    total = int(off[-1])
    scene = O.Scene(z["query"], np.zeros(total, np.float32), np.zeros(total, np.int32), z["model"],
                    np.zeros(total, np.float32), np.zeros(total, np.int32), np.zeros(total, np.int32),
                    np.zeros((1, 2)), np.ones((1, 2)), int(z["width"]), int(z["height"]))
    ids = np.arange(total)
    bins = []
    for b in range(n):
        mem = list(range(int(off[b]), int(off[b + 1])))
        fit = O.affine_fit([tuple(float(v) for v in r) for r in z["model"][mem]],
                           [tuple(float(v) for v in r) for r in z["query"][mem]])
        np.testing.assert_array_equal(np.asarray(fit), z["first_params"][b])
        bn = O.Bin(0, (0, 0, 0, int(z["isigma"][b])), (1500, 1000), mem[0], (0.0, 0.0, 0.0, 1.0))
        bn.members, bn.votes = mem, len(mem)
        bins.append(bn)
    live = O.affine_verify(scene, ids, ids, bins, int(z["threshold"]))
    np.testing.assert_array_equal([any(b is v for v in live) for b in bins], z["live"])
    np.testing.assert_array_equal([b.votes for b in bins], z["votes"])
    np.testing.assert_array_equal(np.array([b.affine for b in bins]), z["last_params"])
    keep = np.zeros(total, bool)
    for b in bins:
        keep[b.members] = True
    np.testing.assert_array_equal(keep, z["keep"])
    assert 0 < z["live"].sum() < n and 0 < z["keep"].sum() < total
