"""DetectionPipeline (array-level API) end to end against the golden fixtures: one call from
descriptors to final poses must give what the reference's Main flow gave on the same scene."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("name", ["scene_single", "scene_multi", "scene_tiny", "scene_c1"])
def test_pipeline_to_final_pose_equals_reference(name):
    from sod_b200.pipeline import DetectionPipeline, ModelDatabase
    z = np.load(GOLD / f"{name}.npz")
    db = ModelDatabase(z["in_m_des"].astype(np.uint8), z["in_m_xy"], z["in_m_angle"], z["in_m_octave"],
                       z["in_m_image"], z["in_img_centroid"], z["in_img_size"])
    nq = len(z["in_q_xy"])
    pipe = DetectionPipeline(db, nq, np.array([[int(z["in_width"]), int(z["in_height"])]], np.int32),
                             bins=int(z["bins"]), vote_threshold=int(z["vote_thr"]),
                             affine_threshold=int(z["affine_thr"]))   # default = the reference's single Hough space
    assert pipe.spaces_per_frame == 1
    out = pipe.detect(z["in_q_des"].astype(np.uint8), z["in_q_xy"], z["in_q_angle"], z["in_q_octave"],
                      np.zeros(nq, np.int32))
    assert out["n_matches"] == len(z["match_q"]) and out["n_near_edge"] == 0
    assert int((out["status"] & 1).sum()) == len(z["live_keys"])
    poses = pipe.final_poses(out)
    got = np.array([[c[0], c[1], o, s, sh[0], sh[1]] for (c, o, s, sh) in poses.get(0, [])], np.float64).reshape(-1, 6)
    assert got.shape == z["final_pose"].shape
    np.testing.assert_allclose(got, z["final_pose"], rtol=1e-10, atol=1e-9)


def test_overlapped_batches_equal_one_at_a_time():
    """detect_batches (copy stream + two query-buffer sets) returns, batch by batch, exactly what
    detect returns for the same batch; the batches differ so that a mixed-up buffer would show."""
    import torch
    import bench
    from sod_b200.pipeline import DetectionPipeline, ModelDatabase
    args = type("A", (), dict(objects=40, kp_per_object=400, frames=4, per_frame=1200, instances=3,
                              inlier_frac=0.2, false_frac=0.02))()
    dev = torch.device("cuda")
    wl = bench.make_workload(args, dev)
    db = bench.make_database(wl)
    nq = args.frames * args.per_frame
    host = [torch.as_tensor(wl[k]).cpu() for k in ("q_des", "q_xy", "q_angle", "q_octave", "q_frame")]
    batches = []
    for shift in (0, 1, 2, 3, 4):           # rotate the query rows of each frame: different matches per batch
        perm = torch.arange(nq).view(args.frames, -1).roll(shift * 37, 1).roll(shift, 0).reshape(-1)
        b = [h[perm].contiguous().pin_memory() for h in host]
        b[4] = host[4].clone().pin_memory()   # frame id stays positional
        batches.append(tuple(b))
    pipe = DetectionPipeline(db, nq, wl["frame_wh"], device=dev)
    want = [pipe.detect(*b) for b in batches]
    got = list(pipe.detect_batches(iter(batches)))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g["n_matches"] == w["n_matches"] and g["n_valid"] == w["n_valid"] and w["n_valid"] > 0
        np.testing.assert_array_equal(g["idx"], w["idx"])
        np.testing.assert_array_equal(g["ok"], w["ok"])
        og, ow = (np.lexsort((x["valid_code"], x["valid_group"])) for x in (g, w))   # compaction order is free
        for k in ("valid_group", "valid_code", "votes", "status", "params"):
            np.testing.assert_array_equal(g[k][og], w[k][ow])
    assert not np.array_equal(want[0]["idx"], want[1]["idx"])


def test_float_descriptors_take_the_bf16_path_through_the_pipeline():
    """Halved (non-integer) descriptors as float32: every product stays exact in the hi/lo split, so
    the whole pipeline must end at the reference's final pose."""
    from sod_b200.pipeline import DetectionPipeline, ModelDatabase
    z = np.load(GOLD / "scene_multi.npz")
    db = ModelDatabase(z["in_m_des"].astype(np.float32) * np.float32(0.5), z["in_m_xy"], z["in_m_angle"],
                       z["in_m_octave"], z["in_m_image"], z["in_img_centroid"], z["in_img_size"])
    nq = len(z["in_q_xy"])
    pipe = DetectionPipeline(db, nq, np.array([[int(z["in_width"]), int(z["in_height"])]], np.int32),
                             per_object_spaces=False)
    assert pipe.float_path
    out = pipe.detect(z["in_q_des"].astype(np.float32) * np.float32(0.5), z["in_q_xy"], z["in_q_angle"],
                      z["in_q_octave"], np.zeros(nq, np.int32))
    assert out["n_matches"] == len(z["match_q"])
    np.testing.assert_array_equal(np.nonzero(out["ok"])[0], z["match_q"])
    np.testing.assert_array_equal(out["idx"][out["ok"].astype(bool), 0], z["match_t"])
    poses = pipe.final_poses(out)
    got = np.array([[c[0], c[1], o, s, sh[0], sh[1]] for (c, o, s, sh) in poses.get(0, [])], np.float64).reshape(-1, 6)
    np.testing.assert_allclose(got, z["final_pose"], rtol=1e-10, atol=1e-9)


def test_small_batch_then_large_batch_on_one_pipeline():
    """A 60-row batch followed by the full batch on the SAME pipeline: the Hough / affine outputs are
    sized once for max_queries, so the second call gets what a fresh pipeline returns (round-1 advisor
    finding: a result object sized by the first call was written past its end by a larger second one)."""
    import torch
    import bench
    from sod_b200.pipeline import DetectionPipeline
    args = type("A", (), dict(objects=40, kp_per_object=400, frames=4, per_frame=1200, instances=3,
                              inlier_frac=0.2, false_frac=0.02))()
    dev = torch.device("cuda")
    wl = bench.make_workload(args, dev)
    db = bench.make_database(wl)
    nq = args.frames * args.per_frame
    full = tuple(torch.as_tensor(wl[k]).cpu() for k in ("q_des", "q_xy", "q_angle", "q_octave", "q_frame"))
    small = tuple(t[:60].contiguous() for t in full)
    pipe = DetectionPipeline(db, nq, wl["frame_wh"], device=dev)
    first = pipe.detect(*small)
    assert first["idx"].shape == (60, 2)
    got = pipe.detect(*full)
    want = DetectionPipeline(db, nq, wl["frame_wh"], device=dev).detect(*full)
    assert want["n_valid"] > 0 and int((want["status"] & 1).sum()) > 0
    np.testing.assert_array_equal(got["idx"], want["idx"])
    og, ow = (np.lexsort((x["valid_code"], x["valid_group"])) for x in (got, want))
    for k in ("valid_group", "valid_code", "votes", "status", "params"):
        np.testing.assert_array_equal(got[k][og], want[k][ow])
    again = pipe.detect(*small)                      # and back down: same as the first small call
    np.testing.assert_array_equal(again["idx"], first["idx"])
    assert again["n_valid"] == first["n_valid"]


def test_affine_result_of_a_smaller_hough_result_is_refused():
    """engine.affine_verify checks the capacity of a caller-held AffineResult; the C ABI flags an
    undersized member_keep (sod_affine_out.cap_votes) instead of writing past it."""
    import torch
    import scenes
    from sod_b200 import engine as E
    sc = scenes.make_scene(seed=11, n_images=1, kp_per_image=800, n_query=600, n_true=200, scales=(1.0,),
                           n_false=50, jitter_frac=0.1, width=1600, height=1200)
    arrays = E.SceneArrays(sc.q_xy, sc.q_angle, sc.q_octave, sc.m_xy, sc.m_angle, sc.m_octave, sc.m_image,
                           sc.img_centroid, sc.img_size.astype(np.float64), np.array([[sc.width, sc.height]], np.int32))
    q = torch.from_numpy(sc.q_des).cuda()
    idx, d2, dist, ok = E.knn_match_ratio(q, E.Matcher(E.prepare_db(torch.from_numpy(sc.m_des).cuda())))
    mq, mt, n_dev = E.compact_matches(idx, ok)
    voter = E.HoughVoter(arrays, 15)
    small_h = voter.vote(mq[:8].contiguous(), mt[:8].contiguous())
    small_a = E.AffineResult(small_h, 5, arrays.device)
    big_h = voter.vote(mq, mt, n_dev)
    with pytest.raises(ValueError):
        E.affine_verify(arrays, mq, mt, big_h, 5, 4, result=small_a)
    # straight through the C ABI with a lying capacity: flagged, not written
    small_a.cap_votes = 4
    small_a.member_keep = torch.zeros(4, dtype=torch.uint8, device=arrays.device)
    guard = torch.full((64,), 7, dtype=torch.uint8, device=arrays.device)    # would be hit by an overrun
    import ctypes as C
    from sod_b200._capi import check, lib
    s, h, o = arrays.struct(), big_h.struct(), small_a.struct()
    check(lib.sod_affine_verify(C.byref(s), E._ptr(mq), E._ptr(mt), C.byref(h), 15, 5, 4, 128.0, 128.0, 0,
                                C.byref(o), E._stream()), "sod_affine_verify")
    torch.cuda.synchronize()
    assert int(small_a.counters[1]) == 1 and int(guard.min()) == 7


def test_graph_replay_equals_eager_detect():
    """detect_replay captures the path for a fixed batch size into one CUDA graph; replays return what the
    eager path returns, also after the query buffers were refilled with another batch."""
    import torch
    import bench
    from sod_b200.pipeline import DetectionPipeline
    args = type("A", (), dict(objects=40, kp_per_object=400, frames=4, per_frame=1200, instances=3,
                              inlier_frac=0.2, false_frac=0.02))()
    dev = torch.device("cuda")
    wl = bench.make_workload(args, dev)
    nq = args.frames * args.per_frame
    pipe = DetectionPipeline(bench.make_database(wl), nq, wl["frame_wh"], device=dev)
    a = tuple(torch.as_tensor(wl[k]).cpu() for k in ("q_des", "q_xy", "q_angle", "q_octave", "q_frame"))
    perm = torch.arange(nq).view(args.frames, -1).roll(91, 1).reshape(-1)
    b = tuple(t[perm].contiguous() if i < 4 else t for i, t in enumerate(a))
    want_a, want_b = pipe.detect(*a), pipe.detect(*b)
    assert not np.array_equal(want_a["idx"], want_b["idx"])
    for batch, want in ((a, want_a), (b, want_b), (a, want_a)):
        pipe.load_queries(*batch)
        got = pipe.fetch(pipe.detect_replay(nq))
        np.testing.assert_array_equal(got["idx"], want["idx"])
        np.testing.assert_array_equal(got["ok"], want["ok"])
        og, ow = (np.lexsort((x["valid_code"], x["valid_group"])) for x in (got, want))
        for k in ("valid_group", "valid_code", "votes", "status", "params"):
            np.testing.assert_array_equal(got[k][og], want[k][ow])
    assert len(pipe._graphs) == 1
