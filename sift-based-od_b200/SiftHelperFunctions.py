"""Keypoint helpers, drop-in for the reference module of the same name (SiftHelperFunctions.py:4-49).

Host-side glue around OpenCV's SIFT keypoints (the input stage stays OpenCV); the octave unpacking
arithmetic itself also lives inside the pose kernel (csrc/sod_hough.cu).
"""
import cv2


def get_centroid(kp):
    """Mean of the keypoint coordinates of one training image (reference :4-14); sums run left to
    right in Python floats so the value is identical to the reference's."""
    sx = 0
    sy = 0
    n = len(kp)
    for k in kp:
        sx = sx + k.pt[0]
        sy = sy + k.pt[1]
    return sx / n, sy / n


def make_temp_kp(kp):
    """cv2.KeyPoint list -> picklable tuples (pt, size, angle, response, octave, class_id) (:16-22)."""
    return [(k.pt, k.size, k.angle, k.response, k.octave, k.class_id) for k in kp]


def unpack_sift_octave(kpt):
    """(octave, layer, scale) from the packed KeyPoint.octave (:25-40): low byte is the octave as
    an 8-bit two's-complement number, next byte the layer, scale = 2^-octave."""
    packed = kpt.octave
    octave = packed & 0xFF
    layer = (packed >> 8) & 0xFF
    if octave >= 128:
        octave -= 256
    scale = float(1 / (1 << octave)) if octave >= 0 else float(1 << -octave)
    return octave, layer, scale


def make_kp(temp_kp):
    """Picklable tuples -> cv2.KeyPoint list (:42-49).  The reference passes `_size=` style keyword
    arguments that OpenCV >= 4.5.3 rejects; positional arguments work on every version."""
    return [cv2.KeyPoint(float(t[0][0]), float(t[0][1]), float(t[1]), float(t[2]), float(t[3]), int(t[4]), int(t[5]))
            for t in temp_kp]
