"""Multi-rank host logic on CPU (gloo, world_size 2): object-aligned sharding, the all-gather of
shard-local top-2 lists and the merge rule reproduce the single-database result bit for bit, and the
per-rank match partition covers every match exactly once.  Kernels are not involved (no GPU here);
the oracle plays the per-shard matcher."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "sift-based-od_b200"))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import sod_oracle as O
    from scenes import sift_like
    from sod_b200.pipeline import shard_bounds

    rng = np.random.default_rng(77)                      # same data on every rank
    n_obj, rows_per = 7, np.array([300, 120, 50, 400, 90, 210, 33])
    db = sift_like(rng, int(rows_per.sum()))
    q = sift_like(rng, 257)
    db[5] = db[900] = q[0]                               # tie across shards
    db[10] = db[11] = q[1]                               # tie inside one shard
    q[2:40] = np.clip(db[rng.integers(0, len(db), 38)].astype(np.int16) + rng.integers(-3, 4, (38, 128)), 0, 255).astype(np.uint8)
    _, _, lo, hi = shard_bounds(n_obj, rows_per, rank, world)
    idx, d2 = O.knn2(q, db[lo:hi])
    idx = np.where(idx >= 0, idx + lo, -1).astype(np.int32)
    mine = torch.from_numpy(np.stack([idx, d2.astype(np.int32)], 0))          # [2, nq, 2]
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    parts = torch.stack(gathered).numpy()                                      # [G, 2, nq, 2]
    gi, gd = O.merge_top2(parts[:, 0], parts[:, 1].astype(np.int64))
    ok = O.ratio_pass(gd, gi)
    keep = ok & (gi[:, 0] >= lo) & (gi[:, 0] < hi)                             # this rank's Hough share
    np.savez(Path(out_dir) / f"rank{rank}.npz", gi=gi, gd=gd, ok=ok, keep=keep, lo=lo, hi=hi)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_top2_merge_equals_single_database(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle import sod_oracle as O
    from scenes import sift_like
    rng = np.random.default_rng(77)
    rows_per = np.array([300, 120, 50, 400, 90, 210, 33])
    db = sift_like(rng, int(rows_per.sum()))
    q = sift_like(rng, 257)
    db[5] = db[900] = q[0]
    db[10] = db[11] = q[1]
    q[2:40] = np.clip(db[rng.integers(0, len(db), 38)].astype(np.int16) + rng.integers(-3, 4, (38, 128)), 0, 255).astype(np.uint8)
    ridx, rd2 = O.knn2(q, db)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    for z in r:                                             # every rank holds the global result
        np.testing.assert_array_equal(z["gi"], ridx)
        np.testing.assert_array_equal(z["gd"], rd2)
    assert ridx[0].tolist() == [5, 900] and ridx[1].tolist() == [10, 11]
    ok = O.ratio_pass(rd2, ridx)
    cover = r[0]["keep"].astype(int) + r[1]["keep"].astype(int)
    np.testing.assert_array_equal(cover, ok.astype(int))    # partition: each match on exactly one rank
    assert int(r[0]["hi"]) == int(r[1]["lo"]) and int(r[0]["lo"]) == 0 and int(r[1]["hi"]) == len(db)


def _np_keys(idx, d2, lo, n_rows):
    """sod_top2_keys in numpy: [n_rows,2] int64, none = INT64_MAX."""
    none = np.iinfo(np.int64).max
    keys = np.full((n_rows, 2), none, np.int64)
    keys[:len(idx)] = np.where(idx >= 0, (d2.astype(np.int64) << 32) | (idx + lo).astype(np.int64), none)
    return keys


def _key_worker(rank, world, port, out_dir):
    """The exchange of the u8 path (include/sod.h, K3 exchange form) with numpy in the kernels' place."""
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import sod_oracle as O
    from scenes import sift_like
    rng = np.random.default_rng(79)
    db = sift_like(rng, 601)
    q = sift_like(rng, 130)                   # 130 rows over 3 ranks: slices of 44 with 2 padding rows
    db[7] = db[420] = db[421] = q[0]          # best and runner-up tie across and inside shards
    db[600] = q[1]                            # the last rank holds ONE row only: it is this row's best
    cuts = [0, 200, 600, 601]
    lo, hi = cuts[rank], cuts[rank + 1]
    idx, d2 = O.knn2(q, db[lo:hi])
    per = (len(q) + world - 1) // world
    keys = torch.from_numpy(_np_keys(idx, d2, lo, per * world))                    # sod_top2_keys
    parts = torch.empty_like(keys)
    dist.all_to_all_single(parts, keys)
    p = np.sort(parts.numpy().reshape(world, per, 2).transpose(1, 0, 2).reshape(per, -1), axis=1)[:, :2]
    mine = torch.from_numpy(np.ascontiguousarray(p))                               # sod_top2_merge_keys
    merged = torch.empty_like(keys)
    dist.all_gather_into_tensor(merged, mine)
    g = merged.numpy()[:len(q)]                                                    # sod_top2_from_keys
    none = np.iinfo(np.int64).max
    gi = np.where(g != none, g & 0xFFFFFFFF, -1).astype(np.int32)
    gd = np.where(g != none, g >> 32, -1)
    np.savez(Path(out_dir) / f"key_rank{rank}.npz", gi=gi, gd=gd)
    dist.barrier()
    dist.destroy_process_group()


def test_key_exchange_is_the_merge(tmp_path):
    """Packed (d2 << 32 | row) keys, all-to-all by query-row slice, two smallest per row, all-gather:
    the single-database top-2 on every rank - ties resolve to the lowest row, a shard with a single
    row takes part, the batch need not divide by the number of ranks."""
    world = 3
    mp.spawn(_key_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle import sod_oracle as O
    from scenes import sift_like
    rng = np.random.default_rng(79)
    db = sift_like(rng, 601)
    q = sift_like(rng, 130)
    db[7] = db[420] = db[421] = q[0]
    db[600] = q[1]
    ridx, rd2 = O.knn2(q, db)
    assert ridx[0].tolist() == [7, 420] and ridx[1, 0] == 600
    for k in range(world):
        z = np.load(tmp_path / f"key_rank{k}.npz")
        np.testing.assert_array_equal(z["gi"], ridx)
        np.testing.assert_array_equal(z["gd"], rd2)


def _pruned_top2(q, shard, lo, thr):
    """What a shard sweep with carried thresholds may return: rows whose d2 - |q|^2 exceeds the row's
    threshold are invisible (the kernel prunes them), the rest compete as usual."""
    from oracle import sod_oracle as O
    qn = (q.astype(np.int64) ** 2).sum(1)
    d = ((q[:, None, :].astype(np.int64) - shard[None].astype(np.int64)) ** 2).sum(2)
    idx = np.full((len(q), 2), -1, np.int32)
    d2 = np.full((len(q), 2), -1, np.int64)
    for r in range(len(q)):
        cand = np.flatnonzero(d[r] - qn[r] <= thr[r])
        order = cand[np.lexsort((cand, d[r, cand]))][:2]
        idx[r, :len(order)] = order + lo
        d2[r, :len(order)] = d[r, order]
    return idx, d2


def _seed_worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "sift-based-od_b200"))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import sod_oracle as O
    from scenes import sift_like
    from sod_b200.pipeline import seed_query_slice, seed_sample_rows, shard_bounds

    rng = np.random.default_rng(78)
    rows_per = np.array([300, 120, 50, 400, 90, 210, 33])
    db = sift_like(rng, int(rows_per.sum()))
    q = sift_like(rng, 700)
    db[5] = db[900] = q[0]
    q[2:60] = np.clip(db[rng.integers(0, len(db), 58)].astype(np.int16) + rng.integers(-3, 4, (58, 128)), 0, 255).astype(np.uint8)
    _, _, lo, hi = shard_bounds(len(rows_per), rows_per, rank, world)
    # seeding: this rank's query slice against the replicated sample -> 2nd-best bound, "none" elsewhere
    none = 0x7F7F7F7F
    thr = torch.full((len(q),), none, dtype=torch.int32)
    sample = db[seed_sample_rows(len(db), 64)]
    s_lo, s_hi = seed_query_slice(len(q), rank, world)
    _, sd2 = O.knn2(q[s_lo:s_hi], sample)
    qn = (q.astype(np.int64) ** 2).sum(1)
    thr[s_lo:s_hi] = torch.from_numpy((sd2[:, 1] - qn[s_lo:s_hi]).astype(np.int32))
    dist.all_reduce(thr, op=dist.ReduceOp.MIN)
    assert int((thr == none).sum()) == 0                                      # the slices tile the batch
    idx, d2 = _pruned_top2(q, db[lo:hi], lo, thr.numpy().astype(np.int64))
    mine = torch.from_numpy(np.stack([idx, d2.astype(np.int32)], 0))
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    parts = torch.stack(gathered).numpy()
    gi, gd = O.merge_top2(parts[:, 0], parts[:, 1].astype(np.int64))
    np.savez(Path(out_dir) / f"seed_rank{rank}.npz", gi=gi, gd=gd, pruned=int((idx < 0).sum()))
    dist.barrier()
    dist.destroy_process_group()


def test_threshold_seeding_keeps_the_sharded_result_exact(tmp_path):
    """Every rank seeds 1/G of the query rows on a replicated sample of the whole database, one
    min-reduce spreads the bounds, the shard sweeps prune with them: the merged top-2 is still the
    single-database result, ties included (host logic of DetectionPipeline.detect_device)."""
    world = 2
    mp.spawn(_seed_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle import sod_oracle as O
    from scenes import sift_like
    rng = np.random.default_rng(78)
    rows_per = np.array([300, 120, 50, 400, 90, 210, 33])
    db = sift_like(rng, int(rows_per.sum()))
    q = sift_like(rng, 700)
    db[5] = db[900] = q[0]
    q[2:60] = np.clip(db[rng.integers(0, len(db), 58)].astype(np.int16) + rng.integers(-3, 4, (58, 128)), 0, 255).astype(np.uint8)
    ridx, rd2 = O.knn2(q, db)
    for k in range(world):
        z = np.load(tmp_path / f"seed_rank{k}.npz")
        np.testing.assert_array_equal(z["gi"], ridx)
        np.testing.assert_array_equal(z["gd"], rd2)
        assert int(z["pruned"]) > 0                          # the bound did hide shard-local candidates


def test_seed_slices_and_sample():
    sys.path.insert(0, str(ROOT / "sift-based-od_b200"))
    from sod_b200.pipeline import QUERY_BLOCK, seed_query_slice, seed_sample_rows
    for n in (0, 1, 255, 256, 257, 5000, 1_280_000):
        for world in (1, 2, 3, 8):
            prev = 0
            for rank in range(world):
                lo, hi = seed_query_slice(n, rank, world)
                assert lo == prev and lo % QUERY_BLOCK == 0 or lo == n
                prev = hi
            assert prev == n
    rows = seed_sample_rows(1_000_000, 16384)
    assert len(rows) == 16384 and rows[0] == 0 and rows[-1] < 1_000_000 and np.all(np.diff(rows) > 0)
    assert seed_sample_rows(10, 64).tolist() == list(range(10))


def test_rank_local_hough_spaces_map_back_to_the_global_numbering():
    """A rank of a database-sharded run numbers only its own objects' Hough spaces; results carry the
    single-GPU ids frame * n_images + object."""
    sys.path.insert(0, str(ROOT / "sift-based-od_b200"))
    from sod_b200.pipeline import global_space_ids, local_hough_spaces, shard_bounds
    n_images, frames = 11, 5
    seen = []
    for world in (1, 2, 3, 8):
        for rank in range(world):
            olo, ohi, _, _ = shard_bounds(n_images, 7, rank, world)
            grp, n_local = local_hough_spaces(n_images, olo, ohi)
            assert n_local == max(ohi - olo, 1) and grp.dtype == np.int32 and grp.min() >= 0 and grp.max() < n_local
            own = np.arange(olo, ohi)
            assert grp[own].tolist() == list(range(len(own)))           # own objects: dense, in order
            f, o = np.meshgrid(np.arange(frames), own, indexing="ij")
            local = (f * n_local + grp[o]).ravel()                      # what the Hough kernel computes
            np.testing.assert_array_equal(global_space_ids(local, n_local, n_images, olo), (f * n_images + o).ravel())
            if world == 1:
                np.testing.assert_array_equal(global_space_ids(local, n_local, n_images, olo), local)
            seen.append((world, rank))
    assert len(seen) == 14


def test_shard_bounds_are_object_aligned_and_cover():
    sys.path.insert(0, str(ROOT / "sift-based-od_b200"))
    from sod_b200.pipeline import shard_bounds
    for world in (1, 2, 3, 4, 8):
        prev = 0
        for rank in range(world):
            olo, ohi, lo, hi = shard_bounds(1000, 1000, rank, world)
            assert lo == prev and lo == olo * 1000 and hi == ohi * 1000
            prev = hi
        assert prev == 1_000_000
    rows = np.array([3, 0, 5, 2])
    assert [shard_bounds(4, rows, r, 2)[2:] for r in range(2)] == [(0, 3), (3, 10)]


def test_single_hough_space_cannot_be_database_sharded():
    """The reference's one Hough space for all model images cannot be split by database rows: votes of
    one space would land on different ranks.  The constructor refuses before touching the device."""
    import pytest
    sys.path.insert(0, str(ROOT / "sift-based-od_b200"))
    from sod_b200.pipeline import DetectionPipeline, ModelDatabase
    n = 40
    db = ModelDatabase(np.zeros((n, 128), np.uint8), np.zeros((n, 2), np.float32), np.zeros(n, np.float32),
                       np.zeros(n, np.int32), np.repeat(np.arange(4, dtype=np.int32), 10), np.zeros((4, 2)),
                       np.ones((4, 2), np.int32))
    with pytest.raises(ValueError, match="per-object"):
        DetectionPipeline(db, 16, np.array([[640, 480]], np.int32), rank=0, world=2)
    with pytest.raises(ValueError, match="per-object"):
        DetectionPipeline(db, 16, np.array([[640, 480]], np.int32), rank=1, world=2, per_object_spaces=False)
    db.object_of_image = np.array([0, 0, 2, 1], np.int32)
    with pytest.raises(ValueError, match="non-decreasing"):
        DetectionPipeline(db, 16, np.array([[640, 480]], np.int32), rank=0, world=2)


def test_shards_split_at_object_boundaries_when_objects_have_several_images():
    sys.path.insert(0, str(ROOT / "sift-based-od_b200"))
    from sod_b200.pipeline import global_space_ids, local_hough_spaces, shard_bounds
    obj_of_img = np.array([0, 0, 0, 1, 2, 2], np.int32)          # 3 objects, 6 training images
    image = np.repeat(np.arange(6), [5, 3, 2, 7, 4, 4])
    rows = np.bincount(obj_of_img[image], minlength=3)
    assert rows.tolist() == [10, 7, 8]
    assert [shard_bounds(3, rows, r, 2) for r in range(2)] == [(0, 1, 0, 10), (1, 3, 10, 25)]
    grp, n_local = local_hough_spaces(obj_of_img, 1, 3)
    assert n_local == 2 and grp.tolist() == [0, 0, 0, 0, 1, 1]
    np.testing.assert_array_equal(global_space_ids(np.array([0, 1, 2, 3]), 2, 3, 1), [1, 2, 4, 5])


def test_rotated_block_order_with_published_thresholds_is_exact():
    """Host logic of sod_match_top2_peer / DetectionPipeline's peer-threshold path, with numpy in the kernel's
    place: G ranks each hold a shard, visit the query blocks in an order rotated by r/G, prune a block with the
    thresholds published so far (the seed of the block's owner, then every earlier visitor's shard-local 2nd
    best) and publish their own when the block is done.  Whatever the interleaving of the ranks, the merged
    top-2 is the single-database result, and later visitors really see tighter bounds."""
    sys.path.insert(0, str(ROOT / "tests"))
    sys.path.insert(0, str(ROOT / "sift-based-od_b200"))
    from oracle import sod_oracle as O
    from scenes import sift_like
    from sod_b200.pipeline import seed_sample_rows
    rng = np.random.default_rng(80)
    world, block, n_blocks = 4, 16, 12
    db = sift_like(rng, 900)
    q = sift_like(rng, block * n_blocks)
    db[17] = db[640] = q[3]                                   # a tie across shards
    q[5:60] = np.clip(db[rng.integers(0, len(db), 55)].astype(np.int16) + rng.integers(-3, 4, (55, 128)), 0, 255).astype(np.uint8)
    cuts = [0, 200, 450, 700, 900]
    qn = (q.astype(np.int64) ** 2).sum(1)
    none = np.int64(0x7F7F7F7F)
    for order_seed in range(3):                               # three different interleavings of the ranks
        thr = np.full(len(q), none)                           # what every rank's array converges to (min-reduce)
        sample = db[seed_sample_rows(len(db), 32)]
        _, sd = O.knn2(q, sample)                             # seeding: 2nd best on the replicated sample
        thr = np.minimum(thr, sd[:, 1] - qn)
        parts_i = np.full((world, len(q), 2), -1, np.int32)
        parts_d = np.full((world, len(q), 2), -1, np.int64)
        events = [(r, k) for r in range(world) for k in range(n_blocks)]
        sched = np.random.default_rng(order_seed)
        pos = [0] * world                                     # every rank walks ITS order; ranks interleave at random
        tighter = 0
        while any(p < n_blocks for p in pos):
            r = int(sched.choice([x for x in range(world) if pos[x] < n_blocks]))
            b = (n_blocks * r // world + pos[r]) % n_blocks   # rotation r / G
            pos[r] += 1
            rows = slice(b * block, (b + 1) * block)
            lo, hi = cuts[r], cuts[r + 1]
            tighter += int((thr[rows] < sd[rows, 1] - qn[rows]).sum())
            idx, d2 = _pruned_top2(q[rows], db[lo:hi], lo, thr[rows])
            parts_i[r, rows], parts_d[r, rows] = idx, d2
            full_i, full_d = O.knn2(q[rows], db[lo:hi])       # what the rank publishes: its shard's own 2nd best
            thr[rows] = np.minimum(thr[rows], np.where(full_i[:, 1] >= 0, full_d[:, 1] - qn[rows], none))
        gi, gd = O.merge_top2(parts_i, parts_d)
        ridx, rd2 = O.knn2(q, db)
        np.testing.assert_array_equal(gi, ridx)
        np.testing.assert_array_equal(gd, rd2)
        assert tighter > 0 and int((parts_i < 0).sum()) > 0   # later visitors had tighter bounds, and they pruned
    assert ridx[3].tolist() == [17, 640]
