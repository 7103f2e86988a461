"""Per-match pose helpers, drop-in for HoughTransformHelperFunctions.py:4-72.

The arithmetic runs in the pose kernel of csrc/sod_hough.cu (fp64, the reference's operation order);
these wrappers evaluate it for a single match / pose so existing callers keep working.  Batched
use goes through Main.apply_hough_transform or sod_b200.dropin.
"""
from SiftHelperFunctions import *  # noqa: F401,F403  (the reference re-exports these)
from sod_b200 import dropin as _dropin


def estimate_object_pose(data):
    """(x, y, alpha, scale_factor): the model centroid of data = (kpM, kpQ, model_size, model_centroid)
    carried into the query image by the keypoint pair's similarity transform (reference :4-37)."""
    pose, _ = _dropin.estimate_poses([data])
    return tuple(float(v) for v in pose[0])


def calculate_bin_index(object_pose, bins, query_image_shape):
    """Base Hough bin (i_x, i_y, i_theta, i_sigma) of a pose; query_image_shape is (H, W, ...)
    (reference :39-72)."""
    return _dropin.pose_bin_indices([object_pose], int(bins), int(query_image_shape[1]),
                                    int(query_image_shape[0]))[0]
