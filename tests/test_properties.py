"""Property tests (hypothesis) of the host logic and of the oracle's own invariants - the same
size-independent properties the GPU tests rely on at full size (shard invariance of the top-2
merge, partition properties of the shard / batch planners, visiting-order recovery)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import sod_oracle as O
from sod_b200.pipeline import shard_bounds
from sod_b200.postprocess import preorder_components
from sod_b200.stream import plan_batches

FAST = settings(max_examples=60, deadline=None)


@FAST
@given(st.lists(st.integers(0, 40), min_size=1, max_size=30), st.integers(1, 9))
def test_shard_bounds_partition_the_rows_object_aligned(rows_per_object, world):
    rows = np.asarray(rows_per_object)
    starts = np.concatenate([[0], np.cumsum(rows)])
    prev_obj, prev_row = 0, 0
    for rank in range(world):
        o_lo, o_hi, r_lo, r_hi = shard_bounds(len(rows), rows, rank, world)
        assert (o_lo, r_lo) == (prev_obj, prev_row) and o_hi >= o_lo
        assert r_lo == starts[o_lo] and r_hi == starts[o_hi]
        prev_obj, prev_row = o_hi, r_hi
    assert prev_obj == len(rows) and prev_row == rows.sum()
    per = int(rows_per_object[0])
    assert shard_bounds(len(rows), per, world - 1, world)[3] == len(rows) * per


@FAST
@given(st.lists(st.integers(0, 50), max_size=40), st.integers(1, 6), st.integers(50, 120))
def test_plan_batches_keeps_order_and_limits(counts, max_frames, max_rows):
    batches = plan_batches(counts, max_frames, max_rows)
    assert [i for b in batches for i in b] == list(range(len(counts)))
    for b in batches:
        assert 1 <= len(b) <= max_frames and sum(counts[i] for i in b) <= max_rows
    for b, nxt in zip(batches, batches[1:]):       # greedy: the next frame did not fit
        assert len(b) == max_frames or sum(counts[i] for i in b) + counts[nxt[0]] > max_rows


@FAST
@given(st.integers(0, 2 ** 32 - 1), st.integers(1, 12), st.integers(2, 60), st.integers(1, 5))
def test_top2_merge_is_shard_invariant(seed, nq, ndb, shards):
    """knn2 on the whole database == merge of knn2 on any contiguous split (ties included: values
    are drawn from a tiny alphabet so that equal distances are common)."""
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 3, (nq, 128)).astype(np.uint8)
    db = rng.integers(0, 3, (ndb, 128)).astype(np.uint8)
    db[rng.integers(0, ndb)] = q[0]
    want_i, want_d = O.knn2(q, db)
    cuts = np.unique(np.concatenate([[0, ndb], rng.integers(0, ndb + 1, shards - 1)]))
    pi, pd = [], []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        i, d = O.knn2(q, db[lo:hi])
        pi.append(np.where(i >= 0, i + lo, -1))
        pd.append(d)
    got_i, got_d = O.merge_top2(np.stack(pi), np.stack(pd))
    np.testing.assert_array_equal(got_i, want_i)
    np.testing.assert_array_equal(got_d, want_d)


@FAST
@given(st.integers(0, 2 ** 32 - 1), st.integers(0, 40), st.floats(0.0, 0.4))
def test_bitset_walk_equals_the_reference_walk(seed, n, density):
    """preorder_components (lowest unvisited neighbour on bit rows) == the reference's recursive dfs
    over ascending neighbour lists, on random graphs."""
    rng = np.random.default_rng(seed)
    adj = np.triu(rng.random((n, n)) < density, 1)
    adj = adj | adj.T
    rows = [int(sum(1 << int(j) for j in np.nonzero(adj[i])[0])) for i in range(n)]
    nbrs = [list(np.nonzero(adj[i])[0]) for i in range(n)]
    assert preorder_components(rows) == O._preorder_components(n, nbrs)


@FAST
@given(st.integers(0, 2 ** 32 - 1), st.integers(1, 30))
def test_ratio_pass_matches_python_float_comparison(seed, n):
    rng = np.random.default_rng(seed)
    d2 = np.sort(rng.integers(0, 400000, (n, 2)), axis=1)
    idx = np.stack([np.arange(n), np.arange(n) + 1], 1).astype(np.int32)
    got = O.ratio_pass(d2, idx)
    want = [float(np.sqrt(np.float32(a))) < 0.75 * float(np.sqrt(np.float32(b))) for a, b in d2]
    assert got.tolist() == want
