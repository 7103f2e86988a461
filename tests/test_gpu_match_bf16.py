"""GPU parity of the bf16 fallback matcher (sod_match_top2_bf16) for non-integer descriptors.

Tolerance (include/sod.h): |d2_gpu - d2_exact| <= tol * (|q|^2 + |t|^2), tol = 2e-5 with the hi/lo
split, 4e-3 for plain bf16.  Integers up to 256 are exact in bf16 and their dot products exact in
fp32, so on integer-valued input the path must agree bit for bit with the exact oracle."""
import numpy as np
import pytest
import torch

from conftest import sift_like
from oracle import sod_oracle as O

pytestmark = pytest.mark.gpu
TOL = {True: 2e-5, False: 4e-3}


def _run(q, db, split=True, index_base=0):
    from sod_b200 import engine as E
    shard = E.prepare_db_float(torch.from_numpy(db).cuda(), index_base, split)
    idx, d2, dist, ok = E.knn_match_ratio_float(torch.from_numpy(q).cuda(), E.FloatMatcher(shard))
    torch.cuda.synchronize()
    return idx.cpu().numpy(), d2.cpu().numpy(), dist.cpu().numpy(), ok.cpu().numpy().astype(bool)


def _exact_d2(q, db, idx):
    diff = q.astype(np.float64)[:, None, :] - db.astype(np.float64)[np.maximum(idx, 0)]
    return (diff * diff).sum(-1)


def _check_tolerance(q, db, split):
    idx, d2, dist, ok = _run(q, db, split)
    ridx, rd2 = O.knn2_float(q, db)
    n = db.shape[0]
    assert (idx[:, 0] >= 0).all() and ((idx[:, 1] >= 0).all() if n >= 2 else (idx[:, 1] == -1).all())
    assert (idx[:, 0] != idx[:, 1]).all()
    qn = (q.astype(np.float64) ** 2).sum(1)
    tn = (db.astype(np.float64) ** 2).sum(1)
    k = 2 if n >= 2 else 1
    scale = qn[:, None] + tn[np.maximum(idx[:, :k], 0)]
    ex = _exact_d2(q, db, idx[:, :k])
    # reported distances are within tolerance of the exact distances of the reported rows
    assert np.all(np.abs(d2[:, :k] - ex) <= TOL[split] * scale + 1e-30), np.abs(d2[:, :k] - ex).max()
    # and the reported rows are the best two up to that tolerance
    slack = 2 * TOL[split] * (qn + tn.max())
    assert np.all(ex[:, 0] <= rd2[:, 0] + slack)
    if k == 2:
        assert np.all(ex[:, 1] <= rd2[:, 1] + slack)
        assert np.all(d2[:, 0] <= d2[:, 1])
    # where the exact gaps are clear of the tolerance the indices are the oracle's
    clear = (rd2[:, 1] - rd2[:, 0] > 2 * slack) if k == 2 else np.ones(len(q), bool)
    assert clear.mean() >= 0.25
    np.testing.assert_array_equal(idx[clear, 0], ridx[clear, 0])
    np.testing.assert_array_equal(dist[:, :k], np.sqrt(d2[:, :k]))
    if k == 2:
        np.testing.assert_array_equal(ok, dist[:, 0].astype(np.float64) < 0.75 * dist[:, 1].astype(np.float64))


def _root_sift(rng, n):
    """RootSIFT-style float descriptors: L1-normalised, square-rooted (unit L2 norm, non-integer)."""
    d = sift_like(rng, n).astype(np.float64) + rng.uniform(0, 1, (n, 128))
    return np.sqrt(d / d.sum(1, keepdims=True)).astype(np.float32)


@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("nq,ndb", [(1, 2), (5, 1), (130, 129), (256, 3000), (1000, 20001)])
def test_rootsift_within_stated_tolerance(nq, ndb, split):
    rng = np.random.default_rng(nq * 7919 + ndb)
    db = _root_sift(rng, ndb)
    q = _root_sift(rng, nq)
    m = min(nq, ndb) // 2           # half of the queries are noisy copies of database rows
    q[:m] = db[rng.permutation(ndb)[:m]] + rng.normal(0, 0.01, (m, 128)).astype(np.float32)
    _check_tolerance(q, db, split)


def test_wide_dynamic_range_split():
    rng = np.random.default_rng(5)
    db = (rng.standard_normal((5000, 128)) * rng.uniform(0.01, 100, (5000, 1))).astype(np.float32)
    q = (rng.standard_normal((300, 128)) * rng.uniform(0.01, 100, (300, 1))).astype(np.float32)
    _check_tolerance(q, db, True)


@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("nq,ndb", [(3, 2), (257, 129), (300, 4097), (64, 50001), (2000, 9000)])
def test_integer_valued_input_is_bit_exact(nq, ndb, split):
    """The kernel mechanics (tiles, segments, column halves, merge, ties) against the exact oracle."""
    rng = np.random.default_rng(nq * 31 + ndb)
    db = sift_like(rng, ndb)
    q = sift_like(rng, nq)
    for r in (0, ndb // 2, ndb - 1):     # duplicates: ties -> lowest index
        db[r] = q[0]
    idx, d2, dist, ok = _run(q.astype(np.float32), db.astype(np.float32), split, index_base=1000)
    ridx, rd2 = O.knn2(q, db)
    np.testing.assert_array_equal(idx, np.where(ridx >= 0, ridx + 1000, -1))
    np.testing.assert_array_equal(d2.astype(np.int64), rd2)
    np.testing.assert_array_equal(dist, O.match_distance(rd2))
    np.testing.assert_array_equal(ok, O.ratio_pass(rd2, ridx))


def test_half_integer_descriptors_scale_exactly():
    """x/2 keeps every product exact: indices equal the integer problem's, d2 is a quarter."""
    rng = np.random.default_rng(9)
    db, q = sift_like(rng, 3000), sift_like(rng, 200)
    idx, d2, _, ok = _run(q.astype(np.float32) * 0.5, db.astype(np.float32) * 0.5)
    ridx, rd2 = O.knn2(q, db)
    np.testing.assert_array_equal(idx, ridx)
    np.testing.assert_array_equal(d2.astype(np.float64) * 4, rd2.astype(np.float64))
    np.testing.assert_array_equal(ok, O.ratio_pass(rd2, ridx))


def test_empty_sides_and_shard_merge():
    from sod_b200 import engine as E
    rng = np.random.default_rng(3)
    q = _root_sift(rng, 70)
    idx, d2, _, ok = _run(q, np.zeros((0, 128), np.float32))
    assert (idx == -1).all() and np.isinf(d2).all() and not ok.any()
    idx, d2, _, ok = _run(np.zeros((0, 128), np.float32), q)
    assert idx.shape == (0, 2)
    # two shards merged = one database
    db = _root_sift(rng, 1500)
    whole = _run(q, db)
    parts = [E.FloatMatcher(E.prepare_db_float(torch.from_numpy(db[lo:hi]).cuda(), lo)).top2(torch.from_numpy(q).cuda())
             for lo, hi in ((0, 700), (700, 1500))]
    gi = torch.stack([p[0] for p in parts]); gd = torch.stack([p[1] for p in parts])
    idx, d2, dist, ok = (t.cpu().numpy() for t in E.merge_top2_float(gi, gd))
    np.testing.assert_array_equal(idx, whole[0])
    np.testing.assert_array_equal(d2, whole[1])
    np.testing.assert_array_equal(ok.astype(bool), whole[3])


def test_non_finite_input_is_rejected():
    from sod_b200 import engine as E
    x = np.ones((4, 128), np.float32); x[2, 5] = np.nan
    with pytest.raises(ValueError, match="NaN"):
        E.prepare_db_float(torch.from_numpy(x).cuda())
