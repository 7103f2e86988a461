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
