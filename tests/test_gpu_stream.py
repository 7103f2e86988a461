"""End to end with real OpenCV SIFT on synthetic images: FrameStream (batched front end) must give,
frame by frame, what the drop-in Main flow gives on the same image, and find the planted objects."""
import cv2
import numpy as np
import pytest

import imaging

pytestmark = pytest.mark.gpu


def _database(rng, n_objects=3):
    from SiftHelperFunctions import get_centroid, make_temp_kp
    from sod_b200.database import PackedDatabase
    rows, objs = [], []
    for i in range(n_objects):
        obj = imaging.textured(rng, 600, 400)
        kp, des = cv2.SIFT_create().detectAndCompute(obj, None)
        rows.append([make_temp_kp(kp), des, (600, 400), get_centroid(kp), f"obj{i}.png"])
        objs.append(obj)
    return PackedDatabase.from_reference_rows(rows), objs


def _main_on(db, feats, kps_db):
    import main as dropin_main
    m = dropin_main.Main()
    m.kp, m.des = kps_db, db.des
    m.img_size_list, m.img_centroid_list = db.per_keypoint_lists()
    m.kp_query = [cv2.KeyPoint(float(feats.xy[i, 0]), float(feats.xy[i, 1]), 1.0, float(feats.angle[i]), 0.0,
                               int(feats.octave[i]), i) for i in range(len(feats))]
    m.des_query = feats.des
    m.rgb_query = np.zeros((feats.size[1], feats.size[0], 3), np.uint8)
    m.image_query_size = feats.size
    m.run_matcher()
    m.apply_hough_transform(15)
    m.get_valid_bins(5)
    m.apply_affine_parameters(4)
    m.post_process()
    return m


def test_stream_equals_main_flow_and_finds_the_objects():
    from sod_b200.pipeline import DetectionPipeline
    from sod_b200.stream import FrameStream, sift_features
    rng = np.random.default_rng(7)
    db, objs = _database(rng)
    placements = [[(0, 0.8, 20, 500, 400)], [(1, 1.0, -35, 700, 450), (2, 0.6, 90, 250, 250)], [],
                  [(2, 1.2, 5, 600, 500)], [(0, 0.7, 170, 800, 300)]]
    frames, truth = [], []
    for pl in placements:
        fr = imaging.textured(rng, 1200, 900, shapes=200)
        t = []
        for (o, s, deg, cx, cy) in pl:
            fr, m = imaging.place(rng, fr, objs[o], s, deg, cx, cy)
            c = db.img_centroid[o]
            t.append((m @ np.array([c[0], c[1], 1.0]), s, deg))
        frames.append(fr)
        truth.append(t)

    pipe = DetectionPipeline(db.to_model_database(), max_queries=20000, frame_wh=np.zeros((2, 2), np.int32),
                             per_object_spaces=False)
    results = dict(FrameStream(pipe, batch_frames=2, workers=4).run(frames))
    assert sorted(results) == list(range(len(frames)))

    kps_db = db.keypoints()
    db_index = {id(k): i for i, k in enumerate(kps_db)}
    for i, fr in enumerate(frames):
        feats = sift_features(fr)
        r = results[i]
        assert r["n_descriptors"] == len(feats)
        m = _main_on(db, feats, kps_db)
        # the matches are the same (query keypoint index, database row) pairs
        # (m.matching_keypoints is consumed by the Hough step only through references, it still holds all)
        assert [t[1].class_id for t in m.matching_keypoints] == r["match_q"].tolist()
        assert [db_index[id(t[0])] for t in m.matching_keypoints] == r["match_t"].tolist()
        assert int(r["live"].sum()) == len(m.valid_bins)
        got = np.array([[c[0], c[1], o, s, sh[0], sh[1]] for (c, o, s, sh) in r["final_pose"]]).reshape(-1, 6)
        want = np.array([[c[0], c[1], o, s, sh[0], sh[1]] for (c, o, s, sh) in m.final_pose]).reshape(-1, 6)
        assert got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9)
        # and every planted object has a final pose at its centroid with its scale
        for (xy, s, deg) in truth[i]:     # position, scale octave and orientation are recovered
            d = np.hypot(got[:, 0] - xy[0], got[:, 1] - xy[1]) if len(got) else np.array([np.inf])
            j = int(d.argmin())
            assert d[j] < 40, (i, xy, got[:, :2])
            assert 0.5 * s <= got[j, 3] <= 2.0 * s
            err = (np.degrees(got[j, 2]) - deg + 180.0) % 360.0 - 180.0
            assert abs(err) < 3.0, (i, deg, np.degrees(got[j, 2]))
        if not truth[i]:
            assert len(got) == 0
