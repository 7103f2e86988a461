"""The drop-in Python surface (main.Main and the helper modules) against the golden fixtures:
same flow, same structures and values as the reference's Main on the same inputs."""
from pathlib import Path

import cv2
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def _kps(xy, size, angle, octave):
    return [cv2.KeyPoint(float(xy[i, 0]), float(xy[i, 1]), float(size[i]), float(angle[i]), 0.0, int(octave[i]), i)
            for i in range(len(xy))]


def _main_for(z):
    import main as dropin_main
    m = dropin_main.Main()
    m.kp = _kps(z["in_m_xy"], z["in_m_size"], z["in_m_angle"], z["in_m_octave"])
    m.des = z["in_m_des"].astype(np.float32)
    m.kp_query = _kps(z["in_q_xy"], z["in_q_size"], z["in_q_angle"], z["in_q_octave"])
    m.des_query = z["in_q_des"].astype(np.float32)
    m.img_size_list = [tuple(int(v) for v in z["in_img_size"][i]) for i in z["in_m_image"]]
    m.img_centroid_list = [tuple(float(v) for v in z["in_img_centroid"][i]) for i in z["in_m_image"]]
    m.rgb_query = np.zeros((int(z["in_height"]), int(z["in_width"]), 3), np.uint8)
    m.image_query_size = (int(z["in_width"]), int(z["in_height"]))
    return m


@pytest.mark.parametrize("name", ["scene_single", "scene_multi", "scene_tiny", "scene_c1"])
def test_main_flow_equals_reference(name):
    z = np.load(GOLD / f"{name}.npz")
    m = _main_for(z)
    m.run_matcher()
    assert [t[1].class_id for t in m.matching_keypoints] == z["match_q"].tolist()
    assert [t[0].class_id for t in m.matching_keypoints] == z["match_t"].tolist()

    m.apply_hough_transform(int(z["bins"]))
    assert [list(k) for k in m.hough_transform.keys()] == z["bin_keys"].tolist()
    bins = list(m.hough_transform.values())
    assert [b.votes for b in bins] == z["bin_votes"].tolist()
    off = z["bin_mem_off"]
    for i, b in enumerate(bins):
        pairs = z["bin_mem"][off[i]:off[i + 1]]
        assert [p[1].class_id for p in b.keypoint_pairs] == pairs[:, 0].tolist()
        assert [p[0].class_id for p in b.keypoint_pairs] == pairs[:, 1].tolist()
    means = np.array([[b.centroid[0], b.centroid[1], b.angle, b.scale, b.img_size[0], b.img_size[1]] for b in bins])
    np.testing.assert_allclose(means, z["bin_means"], rtol=1e-12, atol=1e-9)

    m.get_valid_bins(int(z["vote_thr"]))
    assert [list(b.pose) for b in m.valid_bins] == z["valid_keys"].tolist()
    m.apply_affine_parameters(int(z["affine_thr"]))
    assert [list(b.pose) for b in m.valid_bins] == z["live_keys"].tolist()
    assert [b.votes for b in m.valid_bins] == z["live_votes"].tolist()
    np.testing.assert_allclose(np.array([b.affine_parameters for b in m.valid_bins]).reshape(-1, 6),
                               z["live_params"], rtol=1e-4, atol=1e-6)
    assert len(m.keypoint_pairs) == int(z["n_pairs_after"])

    m.post_process()
    got = np.array([[c[0], c[1], o, s, sh[0], sh[1]] for (c, o, s, sh) in m.final_pose]).reshape(-1, 6)
    np.testing.assert_allclose(got, z["final_pose"], rtol=1e-10, atol=1e-8)


def test_helper_functions_match_oracle():
    from oracle import sod_oracle as O
    import AffineParameters as AP
    import HoughTransformHelperFunctions as HH
    from HoughTransform import perform_hough_transform
    from PoseBin import PoseBin
    z = np.load(GOLD / "scene_single.npz")
    m = _main_for(z)
    m.run_matcher()
    tup = m.matching_keypoints[5]
    pose = HH.estimate_object_pose(tup)
    want = O.estimate_pose(tup[0].pt, tup[0].angle, tup[0].octave, tup[1].pt, tup[1].angle, tup[1].octave, tup[3])
    np.testing.assert_allclose(pose, want, rtol=1e-13, atol=1e-10)
    shape = m.rgb_query.shape
    assert HH.calculate_bin_index(want, 15, shape) == O.bin_index(want, 15, shape[0], shape[1])
    odd = (-50.0, 1e5, 6.2, 3.0)
    assert HH.calculate_bin_index(odd, 15, shape) == O.bin_index(odd, 15, shape[0], shape[1])

    table = perform_hough_transform(m.matching_keypoints, m.rgb_query)
    assert [list(k) for k in table.keys()] == z["bin_keys"].tolist()
    # the legacy signature's separate bin counts (N4): same dict as the oracle's generalisation
    scene = O.Scene(z["in_q_xy"], z["in_q_angle"], z["in_q_octave"], z["in_m_xy"], z["in_m_angle"],
                    z["in_m_octave"], z["in_m_image"], z["in_img_centroid"], z["in_img_size"],
                    int(z["in_width"]), int(z["in_height"]))
    dims = (10, 12, 15, 6)
    legacy = perform_hough_transform(m.matching_keypoints, m.rgb_query, *dims)
    want_tab = O.hough_vote(scene, z["match_q"], z["match_t"], dims)
    assert list(legacy.keys()) == [k[1:] for k in want_tab.keys()]
    assert [b.votes for b in legacy.values()] == [b.votes for b in want_tab.values()]

    # single-bin helpers: one fit, then one residual pass
    big = max(table.values(), key=lambda b: b.votes)
    pairs = list(big.keypoint_pairs)
    pb = PoseBin(big.pose, big.img_size, big.votes, list(pairs), (0, 0, 0, 0))
    AP.AffineParameters(pb)
    mxy = [p[0].pt for p in pairs]
    qxy = [p[1].pt for p in pairs]
    ref = O.affine_fit(mxy, qxy)
    np.testing.assert_allclose(pb.affine_parameters, ref, rtol=1e-6, atol=1e-6)
    x, y = [p[0] for p in mxy], [p[1] for p in mxy]
    got = AP.Calc_x(AP.Gen_A(x, y, len(x)), AP.Gen_b([p[0] for p in qxy], [p[1] for p in qxy], len(x)))
    np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-6)
    _, changed = AP.remove_outliers(pb, m.image_query_size, 128, 128)
    keep = O.affine_residual_keep(ref, mxy, qxy, m.image_query_size[0] * pb.pose[3] / 128,
                                  m.image_query_size[1] * pb.pose[3] / 128)
    assert pb.votes == sum(keep) and changed == (not all(keep))


def test_run_matcher_needs_two_train_rows():
    z = np.load(GOLD / "scene_tiny.npz")
    m = _main_for(z)
    m.des = m.des[:1]
    with pytest.raises(ValueError):
        m.run_matcher()


def test_main_loads_the_packed_database_and_the_reference_pickle(tmp_path):
    """Main.load_database on either file format gives the same matches as setting the fields by hand."""
    from sod_b200.database import PackedDatabase
    z = np.load(GOLD / "scene_multi.npz")
    n = len(z["in_m_xy"])
    db = PackedDatabase(
        des=z["in_m_des"].astype(np.uint8), xy=z["in_m_xy"].astype(np.float32), size=z["in_m_size"].astype(np.float32),
        angle=z["in_m_angle"].astype(np.float32), response=np.zeros(n, np.float32),
        octave=z["in_m_octave"].astype(np.int32), class_id=np.arange(n, dtype=np.int32),
        image=z["in_m_image"].astype(np.int32), img_size=z["in_img_size"].astype(np.int32),
        img_centroid=z["in_img_centroid"].astype(np.float64),
        img_path=np.asarray([f"m{i}.jpg" for i in range(len(z["in_img_size"]))], dtype=str))
    db.save(tmp_path / "db.sodb")
    db.to_pickle(tmp_path / "db.pkl")
    for path in (tmp_path / "db.sodb", tmp_path / "db.pkl"):
        m = _main_for(z)
        m.kp, m.des, m.img_size_list, m.img_centroid_list = [], [], [], []
        m.load_database(path)
        m.run_matcher()
        assert [t[1].class_id for t in m.matching_keypoints] == z["match_q"].tolist()
        assert [t[0].class_id for t in m.matching_keypoints] == z["match_t"].tolist()
        assert [t[2] for t in m.matching_keypoints] == [tuple(int(v) for v in z["in_img_size"][z["in_m_image"][t]])
                                                        for t in z["match_t"]]


def test_main_takes_the_bf16_path_for_non_integer_descriptors():
    """Halved descriptors are non-integer -> bf16 path; every product stays exact, so the matches
    are the golden ones."""
    z = np.load(GOLD / "scene_multi.npz")
    m = _main_for(z)
    m.des = m.des * np.float32(0.5)
    m.des_query = m.des_query * np.float32(0.5)
    m.run_matcher()
    assert [t[1].class_id for t in m.matching_keypoints] == z["match_q"].tolist()
    assert [t[0].class_id for t in m.matching_keypoints] == z["match_t"].tolist()


def test_affine_on_a_bin_list_with_empty_and_tiny_bins():
    """apply_affine_parameters-style use of the drop-in on a caller-built PoseBin list that contains an
    empty bin and single-pair bins (more bins than pairs): the reference simply iterates; the capacity
    of the device output must not be bounded by the number of pairs."""
    from sod_b200 import dropin
    from PoseBin import PoseBin
    kp = lambda x, y: cv2.KeyPoint(float(x), float(y), 1.0, 0.0, 0.0, 0, 0)  # noqa: E731
    rng = np.random.default_rng(5)
    full = PoseBin((3, 4, 5, 2), (100, 100), 0, [], (0, 0, 0, 0))
    for _ in range(9):
        x, y = rng.uniform(0, 500, 2)
        full.keypoint_pairs.append((kp(x, y), kp(2 * x + 10, 2 * y - 5)))
    bins = [PoseBin((0, 0, 0, 2), (100, 100), 0, [], (0, 0, 0, 0)) for _ in range(12)]     # empty bins
    one = PoseBin((1, 1, 1, 2), (100, 100), 0, [(kp(1, 2), kp(3, 4))], (0, 0, 0, 0))
    out = dropin.affine_run(bins[:6] + [full, one] + bins[6:], (1000, 800), 128.0, 128.0, 4, 0)
    assert len(out) == 14
    params, keep, votes, live = out[6]
    assert live and votes == 9 and keep.all()
    np.testing.assert_allclose(params, [2, 0, 0, 2, 10, -5], rtol=1e-6, atol=1e-5)
    assert out[0][0] is None and out[0][2] == 0 and not out[0][3]
    assert out[7][2] <= 1 and not out[7][3]
