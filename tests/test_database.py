"""Packed model database (sod_b200/database.py): lossless both ways against the reference's pickle
rows [temp_kp, des, img_size, centroid, path] (GenerateDatabaseInfo.py:34-37), and the same content
through either file format."""
import pickle

import numpy as np
import pytest

from conftest import sift_like
from sod_b200.database import PackedDatabase


def _rows(rng, counts):
    rows = []
    for i, n in enumerate(counts):
        pts = rng.uniform(0, 1500, (n, 2)).astype(np.float32)
        temp_kp = [((float(pts[k, 0]), float(pts[k, 1])), float(np.float32(rng.uniform(1, 40))),
                    float(np.float32(rng.uniform(0, 360))), float(np.float32(rng.uniform(0, 0.1))),
                    int(rng.integers(0, 8)) | (int(rng.integers(1, 4)) << 8) | (int(rng.integers(0, 2)) * 0xFF), -1)
                   for k in range(n)]
        cent = (sum(p[0][0] for p in temp_kp) / max(n, 1), sum(p[0][1] for p in temp_kp) / max(n, 1))
        rows.append([temp_kp, sift_like(rng, n).astype(np.float32), (1500, 1000 + i), cent, f"img_{i}.jpg"])
    return rows


def _same_rows(a, b):
    assert len(a) == len(b)
    for ra, rb in zip(a, b):
        assert ra[0] == rb[0]
        assert ra[1].dtype == rb[1].dtype and np.array_equal(ra[1], rb[1])
        assert tuple(ra[2]) == tuple(rb[2]) and tuple(ra[3]) == tuple(rb[3]) and ra[4] == rb[4]


def test_rows_round_trip_is_lossless():
    rows = _rows(np.random.default_rng(1), [37, 1, 120])
    db = PackedDatabase.from_reference_rows(rows)
    db.validate()
    assert len(db) == 158 and db.n_images == 3 and db.des.dtype == np.uint8
    assert db.image.tolist() == [0] * 37 + [1] + [2] * 120
    _same_rows(db.to_reference_rows(), rows)


def test_both_file_formats_hold_the_same_database(tmp_path):
    rows = _rows(np.random.default_rng(2), [64, 50])
    with open(tmp_path / "training_data.pkl", "wb") as f:
        pickle.dump(rows, f, pickle.HIGHEST_PROTOCOL)
    a = PackedDatabase.open(tmp_path / "training_data.pkl")
    a.save(tmp_path / "training_data.sodb")
    b = PackedDatabase.open(tmp_path / "training_data.sodb")
    for name in ("des", "xy", "size", "angle", "response", "octave", "class_id", "image", "img_size",
                 "img_centroid", "img_path"):
        x, y = getattr(a, name), getattr(b, name)
        assert x.dtype == y.dtype and np.array_equal(x, y), name
    b.to_pickle(tmp_path / "back.pkl")
    with open(tmp_path / "back.pkl", "rb") as f:
        _same_rows(pickle.load(f), rows)
    sizes, cents = b.per_keypoint_lists()
    assert sizes[0] == (1500, 1000) and sizes[-1] == (1500, 1001) and cents[70] == tuple(rows[1][3])
    kps = b.keypoints()
    assert len(kps) == 114 and kps[5].pt == rows[0][0][5][0] and kps[5].octave == rows[0][0][5][4]


def test_rejects_what_the_u8_path_cannot_hold():
    rows = _rows(np.random.default_rng(3), [8])
    rows[0][1][3, 7] = 0.5
    with pytest.raises(ValueError, match="integer-valued"):
        PackedDatabase.from_reference_rows(rows)
    rows = _rows(np.random.default_rng(3), [8])
    rows[0][1] = rows[0][1][:5]
    with pytest.raises(ValueError, match="does not match"):
        PackedDatabase.from_reference_rows(rows)


def test_empty_database():
    db = PackedDatabase.from_reference_rows([])
    db.validate()
    assert len(db) == 0 and db.n_images == 0 and db.to_reference_rows() == []


def test_build_database_writes_both_formats(tmp_path):
    """GenerateDatabaseInfo.build_database (drop-in for GenerateDatabaseInfo.py:14-37) on images on
    disk: reference row layout in the pickle, the same content in the packed file."""
    import cv2
    import GenerateDatabaseInfo as G
    import imaging
    rng = np.random.default_rng(4)
    img_dir = tmp_path / "train"
    img_dir.mkdir()
    for i, (w, h) in enumerate([(300, 200), (240, 320)]):
        cv2.imwrite(str(img_dir / f"obj{i}.png"), imaging.textured(rng, w, h, shapes=40))
    (img_dir / "notes.txt").write_text("not an image")          # skipped, as cv2.imread returns None
    rows = G.build_database(str(img_dir), str(tmp_path / "training_data.pkl"), str(tmp_path / "training_data.sodb"))
    assert len(rows) == 2
    for temp_kp, des, img_size, centroid, path in rows:
        assert img_size[0] == 1500 and des.shape == (len(temp_kp), 128) and len(temp_kp) > 50   # :23-24 resize
        assert des.dtype == np.float32 and np.array_equal(des, np.rint(des)) and des.max() <= 255
        xs = [t[0][0] for t in temp_kp]
        assert abs(centroid[0] - sum(xs) / len(xs)) < 1e-9 and str(path).endswith(".png")
    with open(tmp_path / "training_data.pkl", "rb") as f:
        _same_rows(pickle.load(f), rows)
    a = PackedDatabase.open(tmp_path / "training_data.pkl")
    b = PackedDatabase.open(tmp_path / "training_data.sodb")
    assert np.array_equal(a.des, b.des) and np.array_equal(a.xy, b.xy) and np.array_equal(a.img_size, b.img_size)
    assert a.n_images == 2 and len(a) == sum(len(r[0]) for r in rows)
