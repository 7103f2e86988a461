"""Packed model database: the array form of what GenerateDatabaseInfo pickles.

The reference stores a pickle: a list of [temp_kp, des, img_size, centroid, path] per training image
(GenerateDatabaseInfo.py:34-37), temp_kp being tuples (pt, size, angle, response, octave, class_id)
(SiftHelperFunctions.py:16-22), and Main.get_query_features flattens it into Python lists of
cv2.KeyPoint on every run (main.py:50-66; broken on OpenCV >= 4.5.3).

PackedDatabase holds the same information as flat numpy arrays - u8 descriptors, structure-of-arrays
keypoints, one row per image for size / centroid / path - so that it loads straight into HBM shards
(`to_model_database`, `DetectionPipeline`) and converts losslessly to and from the reference pickle.
On disk it is a single uncompressed .npz.
"""
from __future__ import annotations

import pickle
from dataclasses import dataclass, fields

import numpy as np

FORMAT_VERSION = 1


@dataclass
class PackedDatabase:
    des: np.ndarray          # u8 [N,128]     SIFT descriptors (integer-valued 0..255)
    xy: np.ndarray           # f32 [N,2]      KeyPoint.pt
    size: np.ndarray         # f32 [N]        KeyPoint.size
    angle: np.ndarray        # f32 [N]        KeyPoint.angle (degrees)
    response: np.ndarray     # f32 [N]        KeyPoint.response
    octave: np.ndarray       # i32 [N]        KeyPoint.octave (packed)
    class_id: np.ndarray     # i32 [N]        KeyPoint.class_id
    image: np.ndarray        # i32 [N]        training image of each keypoint, non-decreasing
    img_size: np.ndarray     # i32 [I,2]      (w,h) per training image
    img_centroid: np.ndarray  # f64 [I,2]     get_centroid(kp) per training image
    img_path: np.ndarray     # str [I]        source file of each training image

    # ------------------------------------------------------------------ reference pickle <-> arrays
    @classmethod
    def from_reference_rows(cls, data) -> "PackedDatabase":
        """data: list of [temp_kp, des, img_size, centroid, path] (the unpickled reference file)."""
        des, xy, size, angle, resp, octv, cid, image = [], [], [], [], [], [], [], []
        img_size, img_cent, img_path = [], [], []
        for i, (temp_kp, d, isz, cent, path) in enumerate(data):
            d = np.asarray(d)
            if d.ndim != 2 or d.shape[1] != 128 or d.shape[0] != len(temp_kp):
                raise ValueError(f"image {i}: descriptor matrix {d.shape} does not match {len(temp_kp)} keypoints")
            if not (np.all(d == np.rint(d)) and d.min(initial=0) >= 0 and d.max(initial=0) <= 255):
                raise ValueError(f"image {i}: descriptors are not integer-valued in 0..255")
            des.append(d.astype(np.uint8))
            for (pt, s, a, r, o, c) in temp_kp:
                xy.append(pt)
                size.append(s)
                angle.append(a)
                resp.append(r)
                octv.append(o)
                cid.append(c)
            image.append(np.full(len(temp_kp), i, np.int32))
            img_size.append(isz)
            img_cent.append(cent)
            img_path.append(str(path))
        cat = lambda xs, dt, shape: (np.concatenate(xs).astype(dt) if xs else np.zeros(shape, dt))  # noqa: E731
        return cls(
            des=cat(des, np.uint8, (0, 128)), xy=np.asarray(xy, np.float32).reshape(-1, 2),
            size=np.asarray(size, np.float32), angle=np.asarray(angle, np.float32),
            response=np.asarray(resp, np.float32), octave=np.asarray(octv, np.int32),
            class_id=np.asarray(cid, np.int32), image=cat(image, np.int32, (0,)),
            img_size=np.asarray(img_size, np.int32).reshape(-1, 2),
            img_centroid=np.asarray(img_cent, np.float64).reshape(-1, 2), img_path=np.asarray(img_path, dtype=str))

    @classmethod
    def from_pickle(cls, path) -> "PackedDatabase":
        with open(path, "rb") as f:
            return cls.from_reference_rows(pickle.load(f))

    def to_reference_rows(self) -> list:
        """Back to the reference's pickle rows (descriptors as float32, as cv2 returns them)."""
        rows = []
        starts = np.searchsorted(self.image, np.arange(len(self.img_size) + 1))
        for i in range(len(self.img_size)):
            lo, hi = int(starts[i]), int(starts[i + 1])
            temp_kp = [((float(self.xy[k, 0]), float(self.xy[k, 1])), float(self.size[k]), float(self.angle[k]),
                        float(self.response[k]), int(self.octave[k]), int(self.class_id[k])) for k in range(lo, hi)]
            rows.append([temp_kp, self.des[lo:hi].astype(np.float32),
                         (int(self.img_size[i, 0]), int(self.img_size[i, 1])),
                         (float(self.img_centroid[i, 0]), float(self.img_centroid[i, 1])), str(self.img_path[i])])
        return rows

    def to_pickle(self, path) -> None:
        with open(path, "wb") as f:
            pickle.dump(self.to_reference_rows(), f, pickle.HIGHEST_PROTOCOL)

    # ------------------------------------------------------------------ packed file
    def save(self, path) -> None:
        arrays = {f.name: getattr(self, f.name) for f in fields(self)}
        with open(path, "wb") as f:
            np.savez(f, format_version=np.int32(FORMAT_VERSION), **arrays)

    @classmethod
    def load(cls, path) -> "PackedDatabase":
        with np.load(path, allow_pickle=False) as z:
            if int(z["format_version"]) != FORMAT_VERSION:
                raise ValueError(f"{path}: unsupported packed database version {int(z['format_version'])}")
            return cls(**{f.name: z[f.name] for f in fields(cls)})

    @classmethod
    def open(cls, path) -> "PackedDatabase":
        """Either format, by content: the reference pickle or the packed .npz."""
        with open(path, "rb") as f:
            magic = f.read(2)
        return cls.load(path) if magic == b"PK" else cls.from_pickle(path)

    # ------------------------------------------------------------------ consumers
    def __len__(self) -> int:
        return int(self.des.shape[0])

    @property
    def n_images(self) -> int:
        return int(self.img_size.shape[0])

    def validate(self) -> None:
        n = len(self)
        for name in ("xy", "size", "angle", "response", "octave", "class_id", "image"):
            if getattr(self, name).shape[0] != n:
                raise ValueError(f"{name} has {getattr(self, name).shape[0]} rows, descriptors {n}")
        if n and (np.any(np.diff(self.image) < 0) or self.image.min() < 0 or self.image.max() >= self.n_images):
            raise ValueError("image ids must be non-decreasing and within range")

    def to_model_database(self):
        """The array bundle sod_b200.pipeline.DetectionPipeline shards across GPUs."""
        from .pipeline import ModelDatabase
        self.validate()
        return ModelDatabase(self.des, self.xy, self.angle, self.octave, self.image, self.img_centroid,
                             self.img_size)

    def keypoints(self):
        """cv2.KeyPoint list for the object-level API (positional arguments: works on every OpenCV)."""
        import cv2
        return [cv2.KeyPoint(float(self.xy[k, 0]), float(self.xy[k, 1]), float(self.size[k]), float(self.angle[k]),
                             float(self.response[k]), int(self.octave[k]), int(self.class_id[k]))
                for k in range(len(self))]

    def per_keypoint_lists(self):
        """(img_size_list, img_centroid_list) as Main holds them: one tuple per keypoint (main.py:54-60)."""
        sizes = [(int(w), int(h)) for w, h in self.img_size]
        cents = [(float(x), float(y)) for x, y in self.img_centroid]
        return [sizes[i] for i in self.image], [cents[i] for i in self.image]
