"""DetectionPipeline (array-level API) end to end against the golden fixtures: one call from
descriptors to final poses must give what the reference's Main flow gave on the same scene."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("name", ["scene_single", "scene_multi", "scene_tiny"])
def test_pipeline_to_final_pose_equals_reference(name):
    from sod_b200.pipeline import DetectionPipeline, ModelDatabase
    z = np.load(GOLD / f"{name}.npz")
    db = ModelDatabase(z["in_m_des"].astype(np.uint8), z["in_m_xy"], z["in_m_angle"], z["in_m_octave"],
                       z["in_m_image"], z["in_img_centroid"], z["in_img_size"])
    nq = len(z["in_q_xy"])
    pipe = DetectionPipeline(db, nq, np.array([[int(z["in_width"]), int(z["in_height"])]], np.int32),
                             bins=int(z["bins"]), vote_threshold=int(z["vote_thr"]),
                             affine_threshold=int(z["affine_thr"]), per_object_spaces=False)
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
    db = ModelDatabase(wl["db_des"], wl["m_xy"], wl["m_angle"], wl["m_octave"], wl["m_image"],
                       wl["img_centroid"], wl["img_size"])
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
