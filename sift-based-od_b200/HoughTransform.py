"""perform_hough_transform, drop-in for the reference's legacy entry point (HoughTransform.py:8-73).

The reference function is dead code that raises IndexError on its first vote; this one
has the semantics of the live path, Main.apply_hough_transform (main.py:89-119), and runs on the GPU.
"""
from PoseBin import *  # noqa: F401,F403
from SiftHelperFunctions import *  # noqa: F401,F403
from PoseBin import PoseBin
from sod_b200 import dropin as _dropin


def perform_hough_transform(matching_keypoints, image_query, bin_x=15, bin_y=15, bin_theta=15, bin_sigma=15):
    """{(i_x, i_y, i_theta, i_sigma): PoseBin} for the matches [(kpM, kpQ, model_size, model_centroid)];
    image_query only needs .shape = (H, W, ...).  The voting kernel uses one bin count for all four
    dimensions, so differing counts are rejected."""
    if not (bin_x == bin_y == bin_theta == bin_sigma):
        raise NotImplementedError("per-dimension bin counts are not supported by the voting kernel yet")
    h, w = int(image_query.shape[0]), int(image_query.shape[1])
    return _dropin.hough_dict(list(matching_keypoints), w, h, int(bin_x), PoseBin)
