"""Launched by tests/test_gpu_multi.py under torchrun: the database-sharded pipeline on G GPUs must
return exactly what the single-GPU pipeline returns (match indices, distances, ratio flags, and the
union over ranks of the verified bins)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "sift-based-od_b200"))

import bench  # noqa: E402
from sod_b200.pipeline import DetectionPipeline  # noqa: E402


def same(a: dict, b: dict, what: str) -> None:
    """Two fetched results agree: match lists row by row, verified bins as sets (compaction order is free)."""
    oa = np.lexsort((a["valid_code"], a["valid_group"]))
    ob = np.lexsort((b["valid_code"], b["valid_group"]))
    for k in ("idx", "ok"):
        assert np.array_equal(a[k], b[k]), f"{what} changed {k}"
    for k in ("valid_group", "valid_code", "votes", "status", "params"):
        assert np.array_equal(a[k][oa], b[k][ob]), f"{what} changed {k}"


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    args = type("A", (), dict(objects=60, kp_per_object=500, frames=6, per_frame=1500, instances=4,
                              inlier_frac=0.2, false_frac=0.02))()
    wl = bench.make_workload(args, dev)
    nq = args.frames * args.per_frame
    db = bench.make_database(wl)
    q_dev = (wl["q_des"], wl["q_xy"], wl["q_angle"], wl["q_octave"], wl["q_frame"])
    q = tuple(torch.as_tensor(a).cpu() for a in q_dev)    # HOST batch, the same on every rank
    # host batch: every rank uploads its slice of the rows, an all-gather replicates it (load_queries)
    pipe = DetectionPipeline(db, nq, wl["frame_wh"], rank=rank, world=world, device=dev)
    assert pipe.seed_matcher is None and pipe.spaces_per_frame == args.objects
    assert pipe.peer is not None, f"peer-memory exchange unavailable: {pipe.peer_error}"
    sharded = pipe.detect(*q)                                 # 9000 rows: the exchange goes through peer memory
    same(sharded, DetectionPipeline(db, nq, wl["frame_wh"], rank=rank, world=world, device=dev,
                                    exchange="scatter").detect(*q), "the NCCL scatter exchange")
    # the whole path as one CUDA-graph replay (latency configuration), twice, and eager again after it
    pipe.load_queries(*q)
    for _ in range(2):
        same(sharded, pipe.fetch(pipe.detect_replay(nq)), "the CUDA-graph replay")
    same(sharded, pipe.detect(*q), "eager after graph")
    # device-resident batch (no slicing), and a host batch uploaded whole by every rank
    same(sharded, pipe.detect(*q_dev), "device-resident inputs")
    same(sharded, DetectionPipeline(db, nq, wl["frame_wh"], rank=rank, world=world, device=dev,
                                    replicated_host=False).detect(*q), "whole-batch upload")
    # the same with threshold seeding forced on (it is automatic only for large databases and batches): a
    # replicated 1024-row sample, each rank seeds 1/G of the query rows, one min-reduce, then the shard sweep
    # thresholds over peer memory (the default where peer memory can be mapped): the seeding sweep and the shard
    # sweep publish every finished block's 2nd best into every rank's array, block order rotated by rank
    seeded_pipe = DetectionPipeline(db, nq, wl["frame_wh"], rank=rank, world=world, device=dev, seed_rows=1024)
    seeded_pipe.seed_min_queries = 0
    assert seeded_pipe.seed_matcher is not None and seeded_pipe.peer_thr is not None, seeded_pipe.peer_error
    seeded = seeded_pipe.detect(*q)
    for _ in range(3):                                       # batch after batch: the two arrays alternate
        same(sharded, seeded_pipe.detect(*q), "peer thresholds, repeated batch")
    # the default of a large database: peer thresholds WITHOUT a seeding sample (forced here: the test database is
    # small), i.e. only the rotation and the publishing of finished blocks
    bare = DetectionPipeline(db, nq, wl["frame_wh"], rank=rank, world=world, device=dev, thresholds="peer", seed_rows=0)
    bare.seed_min_queries = 0
    assert bare.peer_thr is not None and bare.seed_matcher is None
    for _ in range(2):
        same(sharded, bare.detect(*q), "peer thresholds without seeding")
    for stages in (2, 1):                                    # the NCCL form: all-reduce after seeding (+ mid-sweep)
        ar = DetectionPipeline(db, nq, wl["frame_wh"], rank=rank, world=world, device=dev, seed_rows=1024,
                               sweep_stages=stages, thresholds="allreduce")
        ar.seed_min_queries = 0
        assert ar.peer_thr is None and ar.sweep_stages == stages
        same(sharded, ar.detect(*q), f"all-reduce thresholds, {stages} stage(s)")
    # and with the gather form of the exchange (all-gather of the lists + merge on every rank)
    gathered_lists = DetectionPipeline(db, nq, wl["frame_wh"], rank=rank, world=world, device=dev,
                                       exchange="gather").detect(*q)
    same(sharded, seeded, "threshold seeding")
    same(sharded, gathered_lists, "the gather exchange")
    # result_rows="own": every rank reads back only the rows of its slice; together they are the batch
    own = DetectionPipeline(db, nq, wl["frame_wh"], rank=rank, world=world, device=dev, result_rows="own").detect(*q)
    _, lo, hi = pipe.own_rows(nq)
    assert own["row_lo"] == lo and np.array_equal(own["idx"], sharded["idx"][lo:hi]) and \
        np.array_equal(own["ok"], sharded["ok"][lo:hi])
    # a small batch after a large one on the same pipeline (latency configuration: no seeding, same buffers)
    small = pipe.detect(*(a[:1000] for a in q))
    assert np.array_equal(small["idx"], sharded["idx"][:1000])
    keys = np.stack([sharded["valid_group"], sharded["valid_code"], sharded["votes"], sharded["status"] & 1], 1)
    params = sharded["params"]
    gathered_k = [None] * world
    gathered_p = [None] * world
    dist.all_gather_object(gathered_k, keys)
    dist.all_gather_object(gathered_p, params)
    if rank == 0:
        single = DetectionPipeline(db, nq, wl["frame_wh"], rank=0, world=1, device=dev).detect(*q)
        assert single["row_lo"] == 0
        assert np.array_equal(sharded["idx"], single["idx"]), "match indices differ between 1 and G GPUs"
        assert np.array_equal(sharded["ok"], single["ok"])
        k1 = np.stack([single["valid_group"], single["valid_code"], single["votes"], single["status"] & 1], 1)
        kg = np.concatenate(gathered_k)
        pg = np.concatenate(gathered_p)
        o1 = np.lexsort((k1[:, 1], k1[:, 0]))
        og = np.lexsort((kg[:, 1], kg[:, 0]))
        assert np.array_equal(k1[o1], kg[og]), "verified bins differ between 1 and G GPUs"
        assert np.array_equal(single["params"][o1], pg[og]), "affine parameters differ"
        assert single["n_valid"] > 10 and (k1[:, 3] == 1).sum() > 5
        print(f"dist_check ok: world={world} matches={single['n_matches']} valid={single['n_valid']}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
