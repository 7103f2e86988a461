"""FrameStream host logic on the CPU with a stand-in pipeline: batching, order,
ragged and empty frames, look-ahead extraction."""
import threading
import time

import numpy as np
import pytest
import torch

from sod_b200.stream import FrameFeatures, FrameStream, plan_batches


class _Scene:
    def __init__(self, n_frames):
        self.frame_wh = torch.zeros((n_frames, 2), dtype=torch.int32)
        self.groups_per_frame = 1

    @property
    def n_frames(self):
        return self.frame_wh.shape[0]


class _EchoPipeline:
    """Accepts every descriptor whose first byte is odd as a match to database row = second byte."""
    device = torch.device("cpu")

    def __init__(self, max_queries, n_frames):
        self.max_queries, self.scene = max_queries, _Scene(n_frames)
        self.spaces_per_frame = 1
        self.calls = []

    def load_queries(self, des, xy, angle, octave, frame, slot=0, overlap=False):
        self._q = (des.clone(), frame.clone())
        return des.shape[0]

    def detect_device(self, n, slot=0):
        des, frame = self._q
        self.calls.append((n, self.scene.frame_wh.clone()))
        return dict(des=des[:n], frame=frame[:n])

    def fetch(self, r):
        des = r["des"].numpy()
        ok = (des[:, 0] & 1).astype(np.uint8)
        idx = np.stack([des[:, 1].astype(np.int32), np.full(len(des), -1, np.int32)], 1)
        fr = np.unique(r["frame"].numpy())
        return dict(ok=ok, idx=idx, valid_group=fr.astype(np.int32), valid_code=fr.astype(np.int32) * 10,
                    votes=np.full(len(fr), 5, np.int32), status=np.ones(len(fr), np.int32), params=np.zeros((len(fr), 6)))

    def final_poses(self, out):
        return {int(g): [((1.0, 2.0), 0.0, 1.0, (3, 4))] for g in out["valid_group"]}


def _frame(i, n):
    des = np.zeros((n, 128), np.uint8)
    des[:, 0] = np.arange(n) % 256
    des[:, 1] = i
    return FrameFeatures(des, np.zeros((n, 2), np.float32), np.zeros(n, np.float32), np.zeros(n, np.int32),
                         (640 + i, 480))


def test_plan_batches():
    assert plan_batches([5, 5, 5, 5], 3, 100) == [[0, 1, 2], [3]]
    assert plan_batches([60, 50, 40, 10], 8, 100) == [[0], [1, 2, 3]]
    assert plan_batches([], 4, 10) == []
    with pytest.raises(ValueError):
        plan_batches([11], 4, 10)


def test_results_in_order_with_ragged_and_empty_frames():
    counts = [7, 0, 30, 12, 0, 0, 25, 3, 9]
    pipe = _EchoPipeline(max_queries=40, n_frames=3)
    seen = []

    def extract(i):
        seen.append(i)
        time.sleep(0.002 * (i % 3))
        return _frame(i, counts[i])

    got = list(FrameStream(pipe, extract, batch_frames=3, workers=4).run(range(len(counts))))
    assert [i for i, _ in got] == list(range(len(counts)))
    for i, r in got:
        assert r["n_descriptors"] == counts[i]
        np.testing.assert_array_equal(r["match_q"], np.arange(counts[i])[np.arange(counts[i]) % 2 == 1])
        assert (r["match_t"] == i).all()
        if counts[i]:
            assert r["final_pose"] and r["votes"].tolist() == [5] and r["live"].tolist() == [True]
        else:
            assert r["final_pose"] == [] and len(r["votes"]) == 0
    # batches respect both limits and carry the frame sizes of their slots
    assert [n for n, _ in pipe.calls] == [37, 12, 37]      # [7,0,30] [12,0,0] [25,3,9]
    assert pipe.calls[0][1][:, 0].tolist() == [640, 641, 642]
    assert sorted(seen) == list(range(len(counts)))


def test_extraction_runs_ahead_of_the_gpu():
    """While a batch is 'on the GPU' (fetch blocks), workers keep extracting later frames."""
    gate = threading.Event()
    extracted = []

    class Slow(_EchoPipeline):
        def fetch(self, r):
            gate.wait(2.0)
            return super().fetch(r)

    def extract(i):
        extracted.append(i)
        if len(extracted) >= 8:
            gate.set()
        return _frame(i, 4)

    got = list(FrameStream(Slow(16, 2), extract, batch_frames=2, workers=2, prefetch=8).run(range(12)))
    assert len(got) == 12 and gate.is_set()


def test_oversized_frame_is_an_error():
    with pytest.raises(ValueError, match="max_queries"):
        list(FrameStream(_EchoPipeline(10, 2), lambda i: _frame(i, 11), batch_frames=2, workers=1).run(range(2)))
