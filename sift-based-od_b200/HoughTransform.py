"""perform_hough_transform, drop-in for the reference's legacy entry point (HoughTransform.py:8-73).

The reference function is dead code that raises IndexError on its first vote; this one
has the semantics of the live path, Main.apply_hough_transform (main.py:89-119), with the legacy
signature's separate bin counts per dimension, and runs on the GPU.
"""
from PoseBin import *  # noqa: F401,F403
from SiftHelperFunctions import *  # noqa: F401,F403
from PoseBin import PoseBin
from sod_b200 import dropin as _dropin


def perform_hough_transform(matching_keypoints, image_query, bin_x=15, bin_y=15, bin_theta=15, bin_sigma=15):
    """{(i_x, i_y, i_theta, i_sigma): PoseBin} for the matches [(kpM, kpQ, model_size, model_centroid)];
    image_query only needs .shape = (H, W, ...).  One bin count per pose dimension, as the reference
    signature has them; their product is limited to 15^4 counters (one SM's shared memory)."""
    h, w = int(image_query.shape[0]), int(image_query.shape[1])
    return _dropin.hough_dict(list(matching_keypoints), w, h, (int(bin_x), int(bin_y), int(bin_theta), int(bin_sigma)),
                              PoseBin)
