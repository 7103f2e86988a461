"""Array-level public API of the hot path: match -> ratio -> Hough -> affine for a batch of query
frames against a (possibly sharded) model database.  This is what bench.py and multi-GPU callers
use; main.Main is the object-level drop-in on top of the same engine.

Multi-GPU: one process per GPU; the database rows are split contiguously and
object-aligned across ranks, the query batch is replicated, every rank computes shard-local top-2
with global indices, one all-gather of 16 B per query row exchanges them, sod_top2_merge reproduces
the single-GPU result (ties included), and each rank runs Hough + affine for the matches of its own
objects only.  There is no other collective.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import engine as E


def shard_bounds(n_objects: int, rows_per_object: np.ndarray | int, rank: int, world: int) -> tuple[int, int, int, int]:
    """Object-aligned contiguous split -> (obj_lo, obj_hi, row_lo, row_hi) of `rank`."""
    obj_lo = n_objects * rank // world
    obj_hi = n_objects * (rank + 1) // world
    if np.isscalar(rows_per_object):
        return obj_lo, obj_hi, obj_lo * int(rows_per_object), obj_hi * int(rows_per_object)
    starts = np.concatenate([[0], np.cumsum(rows_per_object)])
    return obj_lo, obj_hi, int(starts[obj_lo]), int(starts[obj_hi])


@dataclass
class ModelDatabase:
    """Host-side description of the model database (what GenerateDatabaseInfo pickles, as arrays)."""
    des: np.ndarray | torch.Tensor   # u8 [N,128]
    xy: np.ndarray                   # f32 [N,2]
    angle: np.ndarray                # f32 [N]
    octave: np.ndarray               # i32 [N]
    image: np.ndarray                # i32 [N] model image (= object) of each row, non-decreasing
    img_centroid: np.ndarray         # f64 [n_images,2]
    img_size: np.ndarray             # [n_images,2] (w,h)


class DetectionPipeline:
    def __init__(self, db: ModelDatabase, max_queries: int, frame_wh: np.ndarray, rank: int = 0,
                 world: int = 1, group=None, bins: int = 15, vote_threshold: int = 5,
                 affine_threshold: int = 4, per_object_spaces: bool = True, device: str | torch.device = "cuda",
                 shard: str = "db"):
        """shard="db": database rows split over the ranks, one all-gather of top-2.
        shard="frames": database replicated on every rank, the caller gives each rank its own frames
        (no collective at all); the pipeline then behaves exactly like a single-GPU one."""
        if shard not in ("db", "frames"):
            raise ValueError("shard must be 'db' or 'frames'")
        self.shard_mode = shard
        if shard == "frames":
            rank, world = 0, 1
        self.rank, self.world, self.group = rank, world, group
        self.device = torch.device(device)
        self.bins, self.vote_threshold, self.affine_threshold = bins, vote_threshold, affine_threshold
        n_images = int(db.img_centroid.shape[0])
        image = np.asarray(db.image)
        rows = np.bincount(image, minlength=n_images)
        if np.any(np.diff(image) < 0):
            raise ValueError("database rows must be grouped by model image")
        _, _, self.row_lo, self.row_hi = shard_bounds(n_images, rows, rank, world)
        des = db.des[self.row_lo:self.row_hi]
        des_dev = (des if isinstance(des, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(des)))
        des_dev = des_dev.to(self.device).contiguous()
        # u8 descriptors (what OpenCV SIFT produces) take the exact kind::i8 matcher; float32 descriptors
        # the bf16 path with its stated tolerance (include/sod.h) - queries must then be float32 too.
        self.float_path = des_dev.dtype == torch.float32
        self.seed_tiles = 0
        if self.float_path:
            self.shard = E.prepare_db_float(des_dev, index_base=self.row_lo)
            self.matcher = E.FloatMatcher(self.shard)
        else:
            self.shard = E.prepare_db(des_dev, index_base=self.row_lo)
            self.matcher = E.Matcher(self.shard)
            # database-sharded runs sweep this many stored tiles before the ranks exchange thresholds
            # (the decision must be the same on every rank: it depends on the whole database only)
            n_tiles = self.matcher.n_tiles
            if world > 1 and len(image) // world >= 64 * 128:
                self.seed_tiles = min(max(16, n_tiles // 8), n_tiles)
        self.max_queries = int(max_queries)
        nq = self.max_queries
        dev = self.device
        # query-side device buffers (filled by copy for host inputs, pointers stay stable)
        self.q_des = torch.empty((nq, 128), dtype=torch.float32 if self.float_path else torch.uint8, device=dev)
        self.scene = E.SceneArrays(
            torch.zeros((nq, 2), dtype=torch.float32), torch.zeros(nq, dtype=torch.float32),
            torch.zeros(nq, dtype=torch.int32), db.xy, db.angle, db.octave, db.image, db.img_centroid,
            np.asarray(db.img_size, np.float64), frame_wh, q_frame=torch.zeros(nq, dtype=torch.int32),
            img_group=np.arange(n_images, dtype=np.int32) if per_object_spaces else None,
            groups_per_frame=n_images if per_object_spaces else 1, device=dev)
        self.voter = E.HoughVoter(self.scene, bins)
        self._aff: E.AffineResult | None = None
        # two sets of query-side buffers: set 0 is the one allocated above; set 1 appears on first use
        self._qsets = [dict(des=self.q_des, xy=self.scene.q_xy, angle=self.scene.q_angle,
                            octave=self.scene.q_octave, frame=self.scene.q_frame), None]
        self._copy_stream = None
        self._loaded = [None, None]      # event: the set's host->device copies have landed
        self._consumed = [None, None]    # event: the kernels that read the set have been enqueued and finished
        if world > 1:
            self._gather_idx = torch.empty((world, nq, 2), dtype=torch.int32, device=dev)
            self._gather_d2 = torch.empty((world, nq, 2), dtype=torch.float32 if self.float_path else torch.int32,
                                          device=dev)
        # our kernels per detect_device call; the sample sweep of a database-sharded run adds
        # one match launch and two list merges
        self.launches_per_call = 16 + (3 if self.seed_tiles else 0)

    # ---------------------------------------------------------------- device-resident inputs
    def _qset(self, slot: int) -> dict:
        if self._qsets[slot] is None:
            self._qsets[slot] = {k: torch.empty_like(v) for k, v in self._qsets[0].items()}
        return self._qsets[slot]

    def load_queries(self, des, xy, angle, octave, frame, slot: int = 0, overlap: bool = False) -> int:
        """Copy one batch of query frames (host or device arrays) into query-buffer set `slot`.
        overlap=True issues the copies on the pipeline's copy stream, so that (with pinned host
        memory) they run while the kernels of the other set are busy; detect_device(n, slot) waits
        for them on the device."""
        n = int(des.shape[0])
        if n > self.max_queries:
            raise ValueError("batch larger than max_queries")
        t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(a)  # noqa: E731
        q = self._qset(slot)
        if overlap and self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        stream = self._copy_stream if overlap else torch.cuda.current_stream(self.device)
        if overlap:
            stream.wait_stream(torch.cuda.current_stream(self.device)) if self._consumed[slot] is None else \
                stream.wait_event(self._consumed[slot])
        with torch.cuda.stream(stream):
            for key, src in (("des", des), ("xy", xy), ("angle", angle), ("octave", octave), ("frame", frame)):
                q[key][:n].copy_(t(src), non_blocking=True)
            if overlap:
                self._loaded[slot] = torch.cuda.Event()
                self._loaded[slot].record(stream)
        if not overlap:
            self._loaded[slot] = None
        return n

    def detect_device(self, n: int, slot: int = 0):
        """Run the path on the first n rows of query-buffer set `slot`; everything stays on the device."""
        qs = self._qset(slot)
        if self._loaded[slot] is not None:
            torch.cuda.current_stream(self.device).wait_event(self._loaded[slot])
        sc = self.scene
        self.q_des, sc.q_xy, sc.q_angle, sc.q_octave, sc.q_frame = (qs["des"], qs["xy"], qs["angle"], qs["octave"],
                                                                    qs["frame"])
        q = self.q_des[:n]
        merge = E.merge_top2_float if self.float_path else E.merge_top2
        if self.world > 1 and not self.float_path and self.seed_tiles > 0:
            # Two sweeps per shard with one exchange in between: every rank first sweeps a sample of its
            # shard (the first stored tiles: an even sample of the norm range), the ranks min-reduce the
            # rows' 2nd-best bounds (4 B per query row), and the rest of each shard is swept with the
            # GLOBAL bound from its first tile on.  Without this every shard re-establishes its own
            # thresholds, which costs ~2 ln(n) slow-path updates per row and shard.
            import torch.distributed as dist
            m = self.matcher
            thr = m.new_thresholds(n)
            i1, d1 = m.top2(q, (0, self.seed_tiles), thr)
            dist.all_reduce(thr, op=dist.ReduceOp.MIN, group=self.group)
            i2, d2b = m.top2(q, (self.seed_tiles, m.n_tiles), thr, prepared=True)
            idx, d2 = E.merge_top2(torch.stack([i1, i2]), torch.stack([d1, d2b]))[:2]
        else:
            idx, d2 = self.matcher.top2(q)
        if self.world > 1:
            import torch.distributed as dist
            gi, gd = self._gather_idx[:, :n], self._gather_d2[:, :n]
            if n == self.max_queries:
                dist.all_gather_into_tensor(gi, idx, group=self.group)
                dist.all_gather_into_tensor(gd, d2, group=self.group)
            else:
                gi = torch.empty((self.world, n, 2), dtype=torch.int32, device=self.device)
                gd = torch.empty((self.world, n, 2), dtype=d2.dtype, device=self.device)
                dist.all_gather_into_tensor(gi, idx, group=self.group)
                dist.all_gather_into_tensor(gd, d2, group=self.group)
            idx, d2, dist_f, ok = merge(gi, gd)
        else:
            idx, d2, dist_f, ok = merge(idx[None], d2[None])
        lo, hi = (self.row_lo, self.row_hi) if self.world > 1 else (0, 2 ** 31 - 1)
        mq, mt, n_dev = E.compact_matches(idx, ok, lo, hi)
        hough = self.voter.vote(mq, mt, n_dev, detail_min_count=self.vote_threshold)
        if self._aff is None:
            self._aff = E.AffineResult(hough, self.vote_threshold, self.device)
        aff = E.affine_verify(self.scene, mq, mt, hough, self.vote_threshold, self.affine_threshold,
                              result=self._aff)
        self._consumed[slot] = torch.cuda.Event()
        self._consumed[slot].record()
        return dict(idx=idx, d2=d2, dist=dist_f, ok=ok, match_q=mq, match_t=mt, n_matches=n_dev,
                    hough=hough, affine=aff)

    # ---------------------------------------------------------------- host in, host out
    def detect(self, des, xy, angle, octave, frame) -> dict:
        """Host buffers in (pinned for full copy speed), host results out."""
        n = self.load_queries(des, xy, angle, octave, frame)
        r = self.detect_device(n)
        return self.fetch(r)

    def detect_batches(self, batches):
        """Host batches in, host results out, with the host->device copy of batch i+1 overlapping the
        kernels of batch i (two query-buffer sets, one copy stream).  `batches` yields
        (des, xy, angle, octave, frame) tuples of pinned host arrays; results are yielded in order."""
        it = iter(batches)
        nxt = next(it, None)
        slot = 0
        if nxt is not None:
            n_next = self.load_queries(*nxt, slot=slot, overlap=True)
        while nxt is not None:
            r = self.detect_device(n_next, slot)
            nxt = next(it, None)
            if nxt is not None:
                n_next = self.load_queries(*nxt, slot=slot ^ 1, overlap=True)
            yield self.fetch(r)
            slot ^= 1

    def fetch(self, r: dict) -> dict:
        """Device -> host read of one result: matches, counters, verified bins of this rank."""
        counters = torch.cat([r["n_matches"], r["hough"].counters, r["affine"].counters]).cpu().numpy()
        n_m, n_bins, n_votes, n_edge, ovf, n_valid, ovf2 = (int(v) for v in counters)
        if ovf or ovf2:
            raise RuntimeError("output capacity exceeded")
        a = r["affine"]
        out = dict(
            idx=r["idx"].cpu().numpy(), ok=r["ok"].cpu().numpy(), n_matches=n_m, n_bins=n_bins,
            n_votes=n_votes, n_near_edge=n_edge, n_valid=n_valid,
            valid_bin=a.valid_bin[:n_valid].cpu().numpy(), params=a.params[:n_valid].cpu().numpy(),
            votes=a.votes[:n_valid].cpu().numpy(), status=a.status[:n_valid].cpu().numpy())
        h = r["hough"]
        vb = a.valid_bin[:n_valid].long()
        out["valid_group"] = h.bin_group[vb].cpu().numpy()
        out["valid_code"] = h.bin_code[vb].cpu().numpy()
        out["valid_order"] = h.bin_order[vb].cpu().numpy()
        out["valid_mean"] = h.bin_mean[vb].cpu().numpy()
        return out

    def final_poses(self, out: dict) -> dict[int, list]:
        """Main.post_process (main.py:159-168) for every frame of a fetched result: the bins that
        survived the affine stage are clustered by position and orientation (GPU neighbour graphs,
        sod_b200/postprocess.py) -> {frame: [((cx, cy), orientation, scale, (w, h)), ...]}.  Bins
        enter in the reference's order (Hough insertion order inside a space).  With shard="db"
        this covers the objects of this rank only."""
        from .postprocess import post_process_arrays
        live = (out["status"] & 1).astype(bool)
        group, order, mean = out["valid_group"][live], out["valid_order"][live], out["valid_mean"][live]
        frame = group // self.scene.groups_per_frame
        o = np.lexsort((order, group, frame))
        mean, frame = mean[o], frame[o].astype(np.int32)
        clusters, _, _, final = post_process_arrays(mean[:, 0], mean[:, 1], mean[:, 3], mean[:, 2], mean[:, 4],
                                                    mean[:, 5], segment=frame, device=self.device)
        poses: dict[int, list] = {}
        for cl, pose in zip(clusters, final):
            poses.setdefault(int(frame[cl[0]]), []).append(pose)
        return poses

    @staticmethod
    def fetched_bytes(out: dict) -> int:
        return int(sum(v.nbytes for v in out.values() if isinstance(v, np.ndarray)) + 7 * 4)
