"""GPU parity of the Hough voting and affine verification kernels against the golden fixtures (real
reference outputs) and against the oracle on larger seeded inputs."""
from pathlib import Path

import numpy as np
import pytest
import torch

import scenes
from oracle import sod_oracle as O

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def _scene_arrays(src, prefix="in_", groups=False):
    from sod_b200 import engine as E
    g = (lambda k: src[prefix + k])
    img_group = src["img_group"] if groups else None
    gpf = int(np.max(img_group)) + 1 if groups else 1
    return E.SceneArrays(g("q_xy"), g("q_angle"), g("q_octave"), g("m_xy"), g("m_angle"), g("m_octave"),
                         g("m_image"), g("img_centroid"), np.asarray(g("img_size"), np.float64),
                         np.array([[int(g("width")), int(g("height"))]], np.int32),
                         img_group=img_group, groups_per_frame=gpf)


def _decode(code, bins):
    out = np.zeros((len(code), 4), np.int64)
    c = code.astype(np.int64).copy()
    for d in (3, 2, 1, 0):
        out[:, d] = c % bins
        c //= bins
    return out


@pytest.mark.parametrize("name", ["scene_single", "scene_multi", "scene_tiny"])
def test_hough_matches_reference_golden(name):
    from sod_b200 import engine as E
    z = np.load(GOLD / f"{name}.npz")
    bins = int(z["bins"])
    sc = _scene_arrays(z)
    mq = torch.from_numpy(z["match_q"]).cuda()
    mt = torch.from_numpy(z["match_t"]).cuda()
    res = E.HoughVoter(sc, bins).vote(mq, mt)
    h = res.host()
    assert h["n_near_edge"] == 0
    # bins in reference dict insertion order, votes and members bit-exact
    np.testing.assert_array_equal(_decode(h["code"], bins), z["bin_keys"])
    np.testing.assert_array_equal(h["count"], z["bin_votes"])
    goff = z["bin_mem_off"]
    for i in range(h["n_bins"]):
        mem = h["members"][h["offset"][i]:h["offset"][i] + h["count"][i]]
        pairs = z["bin_mem"][goff[i]:goff[i + 1]]
        np.testing.assert_array_equal(z["match_q"][mem], pairs[:, 0])
        np.testing.assert_array_equal(z["match_t"][mem], pairs[:, 1])
    # running means: same recurrence; cos/sin may differ from libm in the last bit
    np.testing.assert_allclose(h["mean"], z["bin_means"], rtol=1e-12, atol=1e-9)
    # poses against the oracle's libm evaluation
    scene = O.Scene(z["in_q_xy"], z["in_q_angle"], z["in_q_octave"], z["in_m_xy"], z["in_m_angle"],
                    z["in_m_octave"], z["in_m_image"], z["in_img_centroid"], z["in_img_size"],
                    int(z["in_width"]), int(z["in_height"]))
    want = np.array([scene.pose_of(int(a), int(b)) for a, b in zip(z["match_q"], z["match_t"])])
    np.testing.assert_allclose(res.pose[:len(want)].cpu().numpy(), want, rtol=1e-13, atol=1e-10)


@pytest.mark.parametrize("name", ["scene_single", "scene_multi", "scene_tiny"])
def test_affine_matches_reference_golden(name):
    from sod_b200 import engine as E
    z = np.load(GOLD / f"{name}.npz")
    bins = int(z["bins"])
    sc = _scene_arrays(z)
    mq = torch.from_numpy(z["match_q"]).cuda()
    mt = torch.from_numpy(z["match_t"]).cuda()
    res = E.HoughVoter(sc, bins).vote(mq, mt)
    aff = E.affine_verify(sc, mq, mt, res, int(z["vote_thr"]), int(z["affine_thr"]))
    h = res.host()
    a = aff.host(h["n_votes"])
    # valid bins (votes >= threshold) in reference order
    rec_pos = {int(r): i for i, r in enumerate(h["rec"])}
    pos = np.array([rec_pos[int(r)] for r in a["valid_bin"]])
    order = np.argsort(pos)
    keys = _decode(h["code"][pos[order]], bins)
    np.testing.assert_array_equal(keys, z["valid_keys"])
    live = a["live"][order]
    np.testing.assert_array_equal(keys[live], z["live_keys"])
    np.testing.assert_array_equal(a["votes"][order][live], z["live_votes"])
    # affine parameters within 1e-4 relative (north star), in practice ~1e-9
    got = a["params"][order][live]
    np.testing.assert_allclose(got, z["live_params"], rtol=1e-4, atol=1e-6)
    assert np.abs(got - z["live_params"]).max() < 1e-6 * max(1.0, np.abs(z["live_params"]).max())
    # surviving members identical
    goff = z["live_mem_off"]
    li = 0
    for j in order:
        if not a["live"][j]:
            continue
        i = pos[j]
        sl = slice(h["offset"][i], h["offset"][i] + h["count"][i])
        mem = h["members"][sl][a["member_keep"][sl]]
        pairs = z["live_mem"][goff[li]:goff[li + 1]]
        np.testing.assert_array_equal(z["match_q"][mem], pairs[:, 0])
        np.testing.assert_array_equal(z["match_t"][mem], pairs[:, 1])
        li += 1
    assert li == len(z["live_keys"])


def test_multi_group_bin_counts_vs_oracle_stress():
    """SURVEY C5 shape at reduced size: per-object Hough spaces, 90 % outliers; bin -> votes must be
    identical to the oracle's vectorised restatement (itself checked against the scalar one)."""
    from sod_b200 import engine as E
    d = scenes.make_match_stress(103, n_objects=40, per_object=1500)
    bins = 15
    sc = _scene_arrays(d, prefix="", groups=True)
    mq = torch.from_numpy(d["match_q"]).cuda()
    mt = torch.from_numpy(d["match_t"]).cuda()
    res = E.HoughVoter(sc, bins).vote(mq, mt)
    h = res.host()
    osc = O.Scene(d["q_xy"], d["q_angle"], d["q_octave"], d["m_xy"], d["m_angle"], d["m_octave"], d["m_image"],
                  d["img_centroid"], d["img_size"], d["width"], d["height"], d["img_group"])
    _, base = O.hough_base_bins_vectorized(osc, d["match_q"], d["match_t"], bins)
    bb = res.base_bin.cpu().numpy().view(np.uint32)[:len(base)]
    np.testing.assert_array_equal(np.stack([(bb >> s) & 0xFF for s in (0, 8, 16, 24)], 1), base)
    keys, counts = O.vote_counts_vectorized(base, d["img_group"][d["m_image"][d["match_t"]]], bins)
    got = h["group"].astype(np.int64) * bins ** 4 + h["code"]
    o = np.argsort(got)
    np.testing.assert_array_equal(got[o], keys)
    np.testing.assert_array_equal(h["count"][o], counts)
    assert h["n_votes"] == counts.sum()
    # scalar oracle on a sample of groups: members and means
    grp = d["img_group"][d["m_image"][d["match_t"]]]
    sel = np.nonzero(np.isin(grp, [0, 17, 39]))[0]
    table = O.hough_vote(osc, d["match_q"][sel], d["match_t"][sel], bins)
    lut = {(int(g), int(c)): i for i, (g, c) in enumerate(zip(h["group"], h["code"]))}
    for key, b in table.items():
        code = ((key[1] * bins + key[2]) * bins + key[3]) * bins + key[4]
        i = lut[(key[0], code)]
        mem = h["members"][h["offset"][i]:h["offset"][i] + h["count"][i]]
        np.testing.assert_array_equal(mem, sel[b.members])
        np.testing.assert_allclose(h["mean"][i], [b.centroid[0], b.centroid[1], b.angle, b.scale, *b.img_size],
                                   rtol=1e-12, atol=1e-9)


def test_compact_matches_is_stable():
    from sod_b200 import engine as E
    rng = np.random.default_rng(2)
    for nq in (1, 31, 1024, 1025, 70001):
        ok = (rng.random(nq) < 0.3).astype(np.uint8)
        idx = rng.integers(0, 10 ** 6, (nq, 2)).astype(np.int32)
        mq, mt, n = E.compact_matches(torch.from_numpy(idx).cuda(), torch.from_numpy(ok).cuda())
        k = int(n.item())
        want = np.nonzero(ok)[0]
        assert k == len(want)
        np.testing.assert_array_equal(mq[:k].cpu().numpy(), want)
        np.testing.assert_array_equal(mt[:k].cpu().numpy(), idx[want, 0])


def test_bins_above_limit_is_rejected_loudly():
    from sod_b200 import _capi
    from sod_b200 import engine as E
    z = np.load(GOLD / "scene_tiny.npz")
    with pytest.raises(_capi.SodError):
        E.HoughVoter(_scene_arrays(z), 16)


@pytest.mark.parametrize("dims", [(15, 12, 10, 6), (9, 15, 15, 15), (25, 25, 9, 9), (1, 1, 1, 1)])
def test_per_dimension_bin_counts_vs_oracle(dims):
    """sod_hough_vote_dims (legacy perform_hough_transform signature): keys in insertion order,
    votes, members and means against the oracle's per-dimension restatement."""
    from sod_b200 import engine as E
    z = np.load(GOLD / "scene_multi.npz")
    sc = _scene_arrays(z)
    mq = torch.from_numpy(z["match_q"]).cuda()
    mt = torch.from_numpy(z["match_t"]).cuda()
    res = E.HoughVoter(sc, dims).vote(mq, mt)
    h = res.host()
    scene = O.Scene(z["in_q_xy"], z["in_q_angle"], z["in_q_octave"], z["in_m_xy"], z["in_m_angle"],
                    z["in_m_octave"], z["in_m_image"], z["in_img_centroid"], z["in_img_size"],
                    int(z["in_width"]), int(z["in_height"]))
    table = O.hough_vote(scene, z["match_q"], z["match_t"], dims)
    _, by, bt, bs = dims
    code = h["code"].astype(np.int64)
    keys = np.stack([code // (by * bt * bs), code // (bt * bs) % by, code // bs % bt, code % bs], 1)
    assert [tuple(k) for k in keys.tolist()] == [k[1:] for k in table.keys()]
    assert h["count"].tolist() == [b.votes for b in table.values()]
    for i, b in enumerate(table.values()):
        mem = h["members"][h["offset"][i]:h["offset"][i] + h["count"][i]]
        assert mem.tolist() == b.members
    want_mean = np.array([[b.centroid[0], b.centroid[1], b.angle, b.scale] for b in table.values()])
    np.testing.assert_allclose(h["mean"][:, :4], want_mean, rtol=1e-12, atol=1e-9)
    # the affine stage decodes the sigma index with the sigma count
    aff = E.affine_verify(sc, mq, mt, res, 5, 4)
    live = O.affine_verify(scene, z["match_q"], z["match_t"], O.valid_bins(table, 5), 4)
    a = aff.host(h["n_votes"])
    assert int(a["live"].sum()) == len(live)


def test_too_many_counters_is_rejected_loudly():
    from sod_b200 import _capi, engine as E
    z = np.load(GOLD / "scene_tiny.npz")
    with pytest.raises(_capi.SodError):
        E.HoughVoter(_scene_arrays(z), (30, 30, 15, 15))
