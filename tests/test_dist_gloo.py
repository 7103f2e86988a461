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
