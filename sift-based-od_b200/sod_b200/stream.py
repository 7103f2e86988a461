"""Batched multi-frame front end: the caller side of the path.

The reference handles one image per process run: Main.get_query_features (main.py:36-46) reads it,
runs OpenCV SIFT, and the matcher starts when that is done.  For a stream of frames the host-side
SIFT is the slow stage, so FrameStream extracts features of upcoming frames on a pool of host
threads (OpenCV releases the GIL) while the GPU runs match -> ratio -> Hough -> affine on the
previous batch; batches are packed into pinned, double-buffered staging while the GPU is busy, their
host->device copies run on the pipeline's copy stream under the previous batch's kernels, and
results are handed to the caller while the next batch runs.  OpenCV SIFT itself stays the extractor (input
stage, out of scope for the GPU path).
"""
from __future__ import annotations

from collections import deque
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import Callable, Iterable, Iterator

import numpy as np


@dataclass
class FrameFeatures:
    """What the path reads from one query image (main.py:43-46): descriptors + keypoint fields."""
    des: np.ndarray      # u8 [n,128]
    xy: np.ndarray       # f32 [n,2]
    angle: np.ndarray    # f32 [n]
    octave: np.ndarray   # i32 [n]
    size: tuple          # (width, height) of the frame

    def __len__(self) -> int:
        return int(self.des.shape[0])


def sift_features(image, max_keypoints: int | None = None) -> FrameFeatures:
    """cv2.SIFT on one BGR / gray image or file path, as get_query_features does (main.py:40-46).
    A SIFT object is created per call: cv2.SIFT is not safe to share between threads."""
    import cv2
    if isinstance(image, (str, bytes)) or hasattr(image, "__fspath__"):
        img = cv2.imread(str(image))
        if img is None:
            raise FileNotFoundError(str(image))
    else:
        img = np.asarray(image)
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) if img.ndim == 3 else img
    kp, des = cv2.SIFT_create().detectAndCompute(gray, None)
    n = len(kp)
    if n == 0:
        des = np.zeros((0, 128), np.float32)
    if max_keypoints is not None and n > max_keypoints:
        kp, des, n = kp[:max_keypoints], des[:max_keypoints], max_keypoints
    if n and not (np.all(des == np.rint(des)) and des.min() >= 0 and des.max() <= 255):
        raise ValueError("SIFT descriptors are not integer-valued 0..255")
    return FrameFeatures(des.astype(np.uint8), np.array([k.pt for k in kp], np.float32).reshape(-1, 2),
                         np.array([k.angle for k in kp], np.float32), np.array([k.octave for k in kp], np.int32),
                         (int(gray.shape[1]), int(gray.shape[0])))


def plan_batches(counts: Iterable[int], max_frames: int, max_rows: int) -> list[list[int]]:
    """Greedy in-order packing of frames (by descriptor count) into batches of at most max_frames
    frames and max_rows descriptor rows; a frame never straddles batches."""
    batches, cur, rows = [], [], 0
    for i, c in enumerate(counts):
        if c > max_rows:
            raise ValueError(f"frame {i} has {c} descriptors, more than the pipeline's max_queries {max_rows}")
        if cur and (len(cur) == max_frames or rows + c > max_rows):
            batches.append(cur)
            cur, rows = [], 0
        cur.append(i)
        rows += c
    if cur:
        batches.append(cur)
    return batches


class _Staging:
    """One set of pinned host buffers for a batch (allocated lazily; torch only needed on the GPU side)."""

    def __init__(self, max_rows: int, max_frames: int, pin: bool):
        import torch
        mk = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=pin)  # noqa: E731
        self.des = mk((max_rows, 128), torch.uint8)
        self.xy = mk((max_rows, 2), torch.float32)
        self.angle = mk(max_rows, torch.float32)
        self.octave = mk(max_rows, torch.int32)
        self.frame = mk(max_rows, torch.int32)
        self.frame_wh = mk((max_frames, 2), torch.int32)

    def fill(self, feats: list[FrameFeatures]) -> int:
        off = 0
        self.frame_wh.zero_()
        for slot, f in enumerate(feats):
            n = len(f)
            self.des[off:off + n] = _t(f.des)
            self.xy[off:off + n] = _t(f.xy)
            self.angle[off:off + n] = _t(f.angle)
            self.octave[off:off + n] = _t(f.octave)
            self.frame[off:off + n] = slot
            self.frame_wh[slot, 0], self.frame_wh[slot, 1] = f.size
            off += n
        return off


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a))


class FrameStream:
    """Runs a DetectionPipeline over a sequence of frames.

    pipeline      a DetectionPipeline built with max_queries >= the rows of one batch and
                  frame_wh of shape [batch_frames, 2] (its values are overwritten per batch)
    extractor     image -> FrameFeatures (default: sift_features)
    batch_frames  frames per GPU pass
    workers       host threads running the extractor
    prefetch      frames whose extraction may be in flight ahead of the GPU

    run(frames) yields (frame_index, result) in input order; result holds the frame's matches and
    verified bins (see _split) and, with poses=True, the final poses of the frame."""

    def __init__(self, pipeline, extractor: Callable[[object], FrameFeatures] = sift_features, batch_frames: int = 8,
                 workers: int = 8, prefetch: int | None = None, poses: bool = True):
        self.pipeline, self.extractor = pipeline, extractor
        self.batch_frames = int(batch_frames)
        if pipeline.scene.n_frames < self.batch_frames:
            raise ValueError("pipeline.frame_wh must have one row per frame of a batch")
        self.workers = int(workers)
        self.prefetch = int(prefetch) if prefetch is not None else 2 * self.batch_frames + self.workers
        self.poses = poses
        self._staging: list[_Staging] = []

    # -------------------------------------------------------------------------------- GPU side
    def _stage(self, feats: list[FrameFeatures], slot: int) -> int:
        """Pack one batch into the pinned staging set `slot` (host work; the GPU may be busy)."""
        p = self.pipeline
        if not self._staging:
            self._on_gpu = p.device.type == "cuda"
            self._staging = [_Staging(p.max_queries, self.batch_frames, self._on_gpu) for _ in range(2)]
            self._copied = [None, None]
        if self._copied[slot] is not None:
            self._copied[slot].synchronize()      # the copies that read this staging set have finished
        return self._staging[slot].fill(feats)

    def _enqueue(self, slot: int, n: int):
        """Host->device copies (on the pipeline's copy stream, so they overlap the kernels of the
        previous batch) and the kernels of one staged batch, all asynchronous."""
        p, st = self.pipeline, self._staging[slot]
        p.scene.frame_wh[:self.batch_frames].copy_(st.frame_wh, non_blocking=True)
        p.load_queries(st.des[:n], st.xy[:n], st.angle[:n], st.octave[:n], st.frame[:n], slot=slot,
                       overlap=self._on_gpu)
        self._copied[slot] = getattr(p, "_loaded", [None, None])[slot]
        return p.detect_device(n, slot)

    def _finish(self, r, feats: list[FrameFeatures], first_index: int):
        """Device->host read of a batch (synchronises) + final poses, split per frame."""
        if r is None:                                # no descriptors in the whole batch
            z = np.zeros(0, np.int32)
            out = dict(ok=np.zeros(0, np.uint8), idx=np.zeros((0, 2), np.int32), valid_group=z, valid_code=z, votes=z,
                       status=z, params=np.zeros((0, 6)))
            return [(first_index + k, self._split(out, {}, k, feats)) for k in range(len(feats))]
        out = self.pipeline.fetch(r)
        poses = self.pipeline.final_poses(out) if self.poses else {}
        return [(first_index + k, self._split(out, poses, k, feats)) for k in range(len(feats))]

    def _split(self, out: dict, poses: dict, slot: int, feats: list[FrameFeatures]) -> dict:
        """The part of a batch result that belongs to frame `slot`: match pairs as (query keypoint
        index within the frame, database row), verified bins, final poses."""
        lo = sum(len(f) for f in feats[:slot])
        hi = lo + len(feats[slot])
        ok = out["ok"][lo:hi].astype(bool)
        gpf = self.pipeline.spaces_per_frame
        mine = (out["valid_group"] // gpf) == slot
        return dict(match_q=np.nonzero(ok)[0].astype(np.int32), match_t=out["idx"][lo:hi, 0][ok],
                    n_descriptors=hi - lo, valid_group=out["valid_group"][mine] % gpf, valid_code=out["valid_code"][mine],
                    votes=out["votes"][mine], live=(out["status"][mine] & 1).astype(bool), params=out["params"][mine],
                    final_pose=poses.get(slot, []))

    # -------------------------------------------------------------------------------- driver
    def run(self, frames: Iterable) -> Iterator[tuple[int, dict]]:
        p = self.pipeline
        with ThreadPoolExecutor(self.workers) as pool:
            pending: deque = deque()          # futures of extracted frames, input order
            it = iter(frames)
            exhausted = False

            def top_up():
                nonlocal exhausted
                while not exhausted and len(pending) < self.prefetch:
                    try:
                        pending.append(pool.submit(self.extractor, next(it)))
                    except StopIteration:
                        exhausted = True

            def next_batch():
                """Longest in-order prefix of extracted frames that fits one GPU pass."""
                feats, rows = [], 0
                while pending and len(feats) < self.batch_frames:
                    f = pending[0].result()
                    if len(f) > p.max_queries:
                        raise ValueError(f"a frame has {len(f)} descriptors, more than max_queries {p.max_queries}")
                    if feats and rows + len(f) > p.max_queries:
                        break
                    pending.popleft()
                    feats.append(f)
                    rows += len(f)
                    top_up()
                return feats

            top_up()
            # Per iteration: stage batch i on the host while the GPU runs batch i-1, read back i-1
            # (the pipeline's result buffers are single: they must be consumed before the next
            # launch), enqueue i, then hand i-1's frames to the caller while i runs.
            index, slot, prev = 0, 0, None
            while True:
                feats = next_batch()
                n = self._stage(feats, slot) if feats else 0
                results = self._finish(*prev) if prev is not None else []
                prev = (self._enqueue(slot, n) if n else None, feats, index) if feats else None
                yield from results
                if prev is None:
                    break
                index += len(feats)
                slot ^= 1
