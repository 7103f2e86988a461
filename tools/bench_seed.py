"""Dev tool: what threshold seeding buys a database shard (one GPU plays one rank).

A 1M-row database, 1.28M queries; the "shard" is the first `rows` rows.  Variants of the shard sweep:
  plain   no threshold array (single-segment kernel instance)
  fresh   a caller-held threshold array that starts empty (the sharing kernel instance, no seeding)
  seeded  thresholds from a sweep of an even 16k-row sample of the WHOLE database (what the ranks hold
          after the min-reduce in DetectionPipeline.detect_device)
  ideal   thresholds = the final global 2nd best (the floor of any threshold-sharing scheme)
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


def main():
    nq, ndb = 1_280_000, 1_000_000
    shards = [int(a) for a in sys.argv[1:]] or [125_000, 500_000]
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
