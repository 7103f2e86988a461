"""Pose clustering: the GPU neighbour graphs + host visiting-order recovery against
the reference's PostProcessing output (fixture postprocess.npz) and the oracle, all bit-exact."""
import types
from pathlib import Path

import numpy as np
import pytest

from oracle import sod_oracle as O

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def _fixture():
    z = np.load(GOLD / "postprocess.npz")
    return z, [int(v) for v in z["img_w"]], [int(v) for v in z["img_h"]]


def test_arrays_equal_reference_fixture():
    from sod_b200 import postprocess as P
    z, w, h = _fixture()
    clusters, subs, ori, final = P.post_process_arrays(z["cx"], z["cy"], z["scale"], z["angle"], w, h)
    assert [i for c in clusters for i in c] == z["cluster_members"].tolist()
    assert [len(c) for c in clusters] == np.diff(z["cluster_off"]).tolist()
    assert [len(s) for s in subs] == z["subs_per_cluster"].tolist()
    assert [i for s in subs for sub in s for i in sub] == z["sub_members"].tolist()
    np.testing.assert_array_equal(np.array(ori), z["orientation"])
    got = np.array([[c[0], c[1], o, s, sh[0], sh[1]] for (c, o, s, sh) in final], np.float64)
    np.testing.assert_array_equal(got, z["final"])


def test_dropin_functions_equal_reference_fixture():
    import PostProcessing as pp
    z, w, h = _fixture()
    bins = [types.SimpleNamespace(centroid=(float(z["cx"][i]), float(z["cy"][i])), scale=float(z["scale"][i]),
                                  angle=float(z["angle"][i]), img_size=(w[i], h[i]), index=i)
            for i in range(len(w))]
    pose_cluster = pp.group_position(bins)
    ori_cluster = pp.group_orientation(pose_cluster)
    final = pp.get_final_pose(pose_cluster, pp.find_max_orientation(ori_cluster))
    assert [b.index for c in pose_cluster for b in c] == z["cluster_members"].tolist()
    assert [b.index for c in ori_cluster for sub in c for b in sub] == z["sub_members"].tolist()
    got = np.array([[c[0], c[1], o, s, sh[0], sh[1]] for (c, o, s, sh) in final], np.float64)
    np.testing.assert_array_equal(got, z["final"])
    assert pp.group_position([]) == [] and pp.group_orientation([]) == []


def test_cluster_past_the_recursion_limit_equals_oracle():
    """One chain-like cluster of > 1000 bins: the reference's recursive dfs dies here."""
    from sod_b200 import postprocess as P
    rng = np.random.default_rng(8)
    n = 2200
    cx = np.concatenate([np.arange(1500) * 100.0 + rng.uniform(-20, 20, 1500), rng.uniform(0, 150000, n - 1500)])
    cy = np.concatenate([rng.uniform(-30, 30, 1500), rng.uniform(2000, 50000, n - 1500)])
    scale = rng.choice([0.5, 1.0, 2.0], n)
    angle = rng.choice([0.1, 0.11, 0.125, 0.5, 2.0], n) + rng.uniform(-0.004, 0.004, n)
    w, h = [1500] * n, [1000] * n
    perm = rng.permutation(n)
    cx, cy, scale, angle = cx[perm], cy[perm], scale[perm], angle[perm]
    want = O.post_process(cx, cy, scale, angle, w, h)
    assert max(len(c) for c in want[0]) > 1000
    clusters, subs, ori, final = P.post_process_arrays(cx, cy, scale, angle, w, h)
    assert clusters == want[0] and subs == want[1]
    assert ori == want[2]
    assert [(c[0], c[1], o, s, sh[0], sh[1]) for (c, o, s, sh) in final] == want[3]


def test_labels_and_segments():
    from sod_b200 import postprocess as P
    z, w, h = _fixture()
    clusters, label = P.cluster_positions(z["cx"], z["cy"], z["scale"], w, h)
    for cl in clusters:
        assert (label[cl] == min(cl)).all()
    # bins of different segments (frames of a batch) never join
    seg = (np.arange(len(w)) % 2).astype(np.int32)
    clusters2, label2 = P.cluster_positions(z["cx"], z["cy"], z["scale"], w, h, segment=seg)
    for cl in clusters2:
        assert len(set(seg[cl].tolist())) == 1 and (label2[cl] == min(cl)).all()
    for s in (0, 1):
        idx = np.nonzero(seg == s)[0]
        sub, _ = P.cluster_positions(z["cx"][idx], z["cy"][idx], z["scale"][idx], [w[i] for i in idx],
                                     [h[i] for i in idx])
        assert sorted(sorted(idx[c].tolist()) for c in sub) == \
            sorted(sorted(c) for c in clusters2 if seg[c[0]] == s)
