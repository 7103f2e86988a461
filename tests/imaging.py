"""Synthetic images with real SIFT structure for end-to-end tests (no dataset in the repo)."""
import cv2
import numpy as np


def textured(rng, w, h, shapes=120):
    """Smooth noise + random filled shapes: plenty of stable SIFT keypoints."""
    img = cv2.GaussianBlur(rng.integers(0, 256, (h, w), dtype=np.uint8), (0, 0), 3)
    img = cv2.normalize(img, None, 40, 215, cv2.NORM_MINMAX)
    for _ in range(shapes):
        c = int(rng.integers(0, 256))
        x, y = int(rng.integers(0, w)), int(rng.integers(0, h))
        if rng.random() < 0.5:
            cv2.circle(img, (x, y), int(rng.integers(4, 28)), c, -1)
        else:
            cv2.rectangle(img, (x, y), (x + int(rng.integers(6, 50)), y + int(rng.integers(6, 50))), c, -1)
    return cv2.GaussianBlur(img, (0, 0), 1.0)


def place(rng, frame, obj, scale, degrees, cx, cy):
    """Paste obj into frame under a similarity transform; returns (frame, 2x3 matrix obj -> frame)."""
    h, w = obj.shape[:2]
    m = cv2.getRotationMatrix2D((w / 2, h / 2), -degrees, scale)
    m[0, 2] += cx - w / 2
    m[1, 2] += cy - h / 2
    size = (frame.shape[1], frame.shape[0])
    warped = cv2.warpAffine(obj, m, size, flags=cv2.INTER_LINEAR)
    mask = cv2.warpAffine(np.full(obj.shape[:2], 255, np.uint8), m, size, flags=cv2.INTER_NEAREST)
    out = frame.copy()
    out[mask > 0] = warped[mask > 0]
    return out, m
