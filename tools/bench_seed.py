"""Dev tool: what threshold seeding buys a database shard (one GPU plays one rank).

A 1M-row database, 1.28M queries; the "shard" is the first `rows` rows.  Variants of the shard sweep:
  plain   no threshold array (single-segment kernel instance)
  fresh   a caller-held threshold array that starts empty (the sharing kernel instance, no seeding)
  seeded  thresholds from a sweep of an even 16k-row sample of the WHOLE database (what the ranks hold
          after the min-reduce in DetectionPipeline.detect_device)
  ideal   thresholds = the final global 2nd best (the floor of any threshold-sharing scheme)
  staged  (with `--staged G`, shards of ndb / G rows) the seeded sweep in two halves with a MIN all-reduce of
          the thresholds between them, the other G - 1 ranks played by this GPU outside the timed regions:
          after half of every shard the bound is the best 2nd best any rank has seen in 1/2 of the database
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "sift-based-od_b200"))
sys.path.insert(0, str(ROOT / "tools"))
from bench_match import sift_like_gpu  # noqa: E402
from sod_b200 import engine as E  # noqa: E402
from sod_b200.pipeline import seed_sample_rows  # noqa: E402


def timed(f, iters=4):
    f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        f()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def staged(db, q, seeded, world, nq, stages=2):
    """Rank 0's shard sweep in `stages` ranges with a min-reduce of all ranks' thresholds between them."""
    rows = db.shape[0] // world
    ms = [E.Matcher(E.prepare_db(db[r * rows:(r + 1) * rows].contiguous(), index_base=r * rows)) for r in range(world)]
    tiles = ms[0].n_tiles
    cuts = [tiles * k // stages for k in range(stages + 1)]
    thr = [seeded.clone() for _ in range(world)]
    total, lists = 0.0, []
    for k in range(stages):
        rng = (cuts[k], cuts[k + 1])
        keep = thr[0].clone()
        total += timed(lambda: ms[0].top2(q, rng, thr[0].copy_(keep)))
        lists.append(ms[0].top2(q, rng, thr[0].copy_(keep)))
        for r in range(1, world):
            ms[r].top2(q, rng, thr[r])
        red = torch.stack(thr).min(0).values          # the MIN all-reduce
        for r in range(world):
            thr[r].copy_(red)
    return total, lists, ms[0]


def main():
    nq, ndb = 1_280_000, 1_000_000
    argv = sys.argv[1:]
    world = 0
    if "--staged" in argv:
        i = argv.index("--staged")
        world = int(argv[i + 1])
        del argv[i:i + 2]
    shards = [int(a) for a in argv] or [125_000, 500_000]
    g = torch.Generator(device="cuda").manual_seed(3)
    db, q = sift_like_gpu(ndb, g), sift_like_gpu(nq, g)
    full = E.Matcher(E.prepare_db(db))
    _, full_d = full.top2(q)
    qn = (q.int() ** 2).sum(1)
    ideal = full.new_thresholds(nq)
    ideal[:nq] = full_d[:, 1] - qn
    seed_m = E.Matcher(E.prepare_db(db[torch.from_numpy(seed_sample_rows(ndb)).cuda()].contiguous()))
    seeded = full.new_thresholds(nq)
    t_seed_all = timed(lambda: seed_m.top2(q, None, seeded.fill_(0x7F7F7F7F)))
    print(f"seed sweep of ALL {nq} rows on {seed_m.shard.n} sample rows: {t_seed_all:.3f} ms "
          f"(a rank does 1/G of it)", flush=True)
    if world:
        for stages in (2, 3, 4):
            t, lists, m0 = staged(db, q, seeded, world, nq, stages)
            # exactness: entries within the final global bound survive in the union of the ranges' lists
            ref_i, ref_d = m0.top2(q)
            mi, md, _, _ = E.merge_top2(torch.stack([l[0] for l in lists]), torch.stack([l[1] for l in lists]))
            keep = (ref_d >= 0) & (ref_d - qn[:, None] <= ideal[:nq, None])
            assert torch.equal(mi[keep], ref_i[keep]) and torch.equal(md[keep], ref_d[keep])
            print(f"staged x{stages} over {world} ranks, shard {ndb // world} rows: {t:8.3f} ms "
                  f"(sum of the range sweeps; + {stages - 1} MIN all-reduce(s) of {nq * 4 / 1e6:.1f} MB)", flush=True)
    for rows in shards:
        m = E.Matcher(E.prepare_db(db[:rows].contiguous()))
        ref_i, ref_d = m.top2(q)
        res = {}
        res["plain"] = timed(lambda: m.top2(q))
        res["fresh"] = timed(lambda: m.top2(q, None, m.new_thresholds(nq)))
        res["seeded"] = timed(lambda: m.top2(q, None, seeded.clone()))
        res["ideal"] = timed(lambda: m.top2(q, None, ideal.clone()))
        # exactness of the seeded sweep: every entry of the shard's own top-2 that is within the global
        # bound (only those can matter after the merge) must come back unchanged
        si, sd = m.top2(q, None, seeded.clone())
        keep = (ref_d >= 0) & (ref_d - qn[:, None] <= seeded[:nq, None])
        assert torch.equal(si[keep], ref_i[keep]) and torch.equal(sd[keep], ref_d[keep])
        print(f"shard {rows:>8} rows: " + "  ".join(f"{k} {v:8.3f} ms" for k, v in res.items()), flush=True)


if __name__ == "__main__":
    main()
